cd /root/repo
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_patch_gemm.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -8
for f in 0 1; do ACX_GEMM_FUSE_REDUCE=$f python tools/update_time.py 200; done
for c in f2 f3 d2 d3; do python tools/conv_one.py $c 3 50; done
python tools/profile_update.py > gpurun_out/run3_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/run3_launches.csv python tools/profile_update.py > gpurun_out/run3_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/run3_launches.csv > gpurun_out/run3_launches.txt 2>&1; tail -40 gpurun_out/run3_launches.txt
