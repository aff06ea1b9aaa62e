cd /root/repo
ACX_GEMM_TRACE=1 python tools/gemm_reduce_probe.py 2>&1 | tail -30
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -4
for f in 0 1; do ACX_GEMM_FUSE_REDUCE=$f python tools/update_time.py 200; done
