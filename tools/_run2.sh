cd /root/repo
for cut in 7,7,2 7,1,18 1,1,128 1,7,18 7,1,16 7,1,9; do ACX_CONV_CUT=$cut python tools/conv_one.py f3 3 50 | sed "s/^/cut $cut: /"; done
for cut in 9,9,1 9,1,14 3,3,14 1,1,128 9,3,4; do ACX_CONV_CUT=$cut python tools/conv_one.py f2 3 50 | sed "s/^/cut $cut: /"; done
for cut in 9,9,1 9,1,14 3,3,14 1,1,128 9,3,4; do ACX_CONV_CUT=$cut python tools/conv_one.py d3 3 50 | sed "s/^/cut $cut: /"; done
for cut in 10,10,1 2,2,32 10,1,12 1,1,128 5,5,5 10,2,6; do ACX_CONV_CUT=$cut python tools/conv_one.py d2 3 50 | sed "s/^/cut $cut: /"; done
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_patch_gemm.py tests/test_gpu_conv.py tests/test_gpu_learner.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -8
for f in 0 1; do ACX_GEMM_FUSE_REDUCE=$f python tools/update_time.py 200; done
