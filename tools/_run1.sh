cd /root/repo
for c in f2 f3 d2 d3; do for pk in 0 1; do ACX_CONV_PACK=$pk python tools/conv_one.py $c 3 50; done; done 2>&1 | grep -v "^$"
python -m pytest tests/test_gpu_conv.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -5
for pk in 0 1; do ACX_CONV_PACK=$pk python tools/update_time.py 200; done
