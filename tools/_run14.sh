cd /root/repo
for s in syrk_conv2 syrk_conv1 wgrad_conv1 fwd_conv2 fc4; do ACX_GEMM_TRACE=1 python tools/gemm_one.py $s 2>&1 | tail -2; done
for c in f2 f3 d2 d3; do ACX_CONV_DEBUG=32 python tools/conv_one.py $c 3 20 2>&1 | tail -2; done
