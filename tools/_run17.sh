cd /root/repo
python -m pytest tests/test_gpu_gather_gemm.py -x -q -m gpu 2>&1 | tail -3
for d in 32 33 34 36 40; do ACX_CONV_DEBUG=$d python tools/conv_one.py f1 3 30 2>&1 | tail -2; done
