cd /root/repo
for i in 1 2; do
for f in 0 1; do for g in 0 1; do ACX_GEMM_FUSE_REDUCE=$f ACX_GATHER=$g python tools/update_time.py 300; done; done
done
export ACX_GEMM_FUSE_REDUCE=0
python tools/profile_update.py > gpurun_out/run6_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/run6_launches.csv python tools/profile_update.py > gpurun_out/run6_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/run6_launches.csv -v > gpurun_out/run6_launches.txt 2>&1; cat gpurun_out/run6_launches.txt
