"""Time acx_gemm shapes under the three split-K reduction modes (acx_debug_set_fuse_reduce)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from actorcritic_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
CASES = [("heads_A 513 sym", 513, 513, 640, True, True, 0), ("fc4_G 512 sym", 512, 512, 640, True, True, 0),
         ("fc4_A 1569 sym", 1569, 1569, 640, True, True, 0), ("fc4 fwd", 672, 512, 1568, False, False, 0),
         ("sym 512 k=5184", 512, 512, 5184, True, True, 0), ("nonsym 512 s4", 512, 512, 2048, False, False, 4)]
for name, m, n, k, trans, sym, splits in CASES:
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.randn((k, m) if trans else (m, k), device="cuda", generator=g)
    a = ops.split_planes(xa, 2)
    b = a if sym else ops.split_planes(torch.randn((k, n) if trans else (n, k), device="cuda", generator=g), 2)
    for mode in (0, 1, 2):
        lib.acx_debug_set_fuse_reduce(mode)
        run = lambda: ops.gemm(a, b, m, n, k, trans=trans, pairs=ops.PAIRS[3], symmetric=sym, splits=splits)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        import ctypes
        ms = ctypes.c_float()
        lib.acx_gemm_enable_timing(1)
        run()
        lib.acx_gemm_last_ms(ctypes.byref(ms))
        lib.acx_gemm_enable_timing(0)
        arr = (ctypes.c_longlong * 12)()
        lib.acx_debug_gemm_trace(arr)
        t = list(arr)
        print("%-18s mode %d: %.1f us per call (incl. torch allocs); kernel alone %.1f us; trace (cycles since fence start) %s" %
              (name, mode, 1e3 * e0.elapsed_time(e1) / 20, 1e3 * ms.value, [x - t[4] for x in t[5:]]))
