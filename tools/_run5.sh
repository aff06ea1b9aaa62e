cd /root/repo
python -m pytest tests/test_gpu_gather_gemm.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_gpu_learner.py tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -15
for f in 0 1; do ACX_GATHER=$f python tools/update_time.py 200; done
ACX_GEMM_FUSE_REDUCE=0 python tools/update_time.py 200
