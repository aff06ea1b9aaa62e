cd /root/repo
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_gather_gemm.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -3
for c in f1 f2 f3 d2 d3; do ACX_CONV_DEBUG=32 python tools/conv_one.py $c 3 30 2>&1 | tail -2; done
python tools/gather_probe.py 2>&1 | tail -6
for i in 1 2; do python tools/update_time.py 300; done
