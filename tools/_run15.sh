cd /root/repo
for i in 1 2; do for f in 0 1 2; do ACX_ACT_PDL=$f python tools/rollout_time.py; done; done
python -m pytest tests/test_gpu_api.py tests/test_gpu_preprocess.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -3
python tools/update_time.py 300
