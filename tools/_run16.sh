cd /root/repo
python -m pytest tests/test_gpu_learner.py tests/test_gpu_conv.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do for f in 0 1; do ACX_CONV1_BRES=$f python tools/update_time.py 300; done; done
