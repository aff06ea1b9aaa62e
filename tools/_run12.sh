cd /root/repo
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_gather_gemm.py tests/test_gpu_patch_gemm.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -4
python tools/gather_probe.py 2>&1 | tail -6
for i in 1 2; do for g in 0 1; do ACX_SYRK_DIAG=$g python tools/update_time.py 300; done; done
