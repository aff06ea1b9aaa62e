cd /root/repo
for i in 1 2; do for g in 0 1 7; do ACX_GATHER=$g python tools/update_time.py 300; done; done
ACX_GATHER=1 python -m pytest tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -3
