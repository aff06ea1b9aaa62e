"""Kernel-only times (library events, warm) of the conv factor / weight-gradient GEMMs on materialised patch matrices
against the same products with operands read in place (acx_gather_t).  usage: gather_probe.py"""
import ctypes
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from actorcritic_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
N = 640


def timed(fn, reps=8):
    lib.acx_gemm_enable_timing(1)
    d = []
    for i in range(reps):
        fn()
        ms = ctypes.c_float(0)
        _lib.check(lib.acx_gemm_last_ms(ctypes.byref(ms)))
        if i >= 3:
            d.append(ms.value * 1e3)
    lib.acx_gemm_enable_timing(0)
    return sum(d) / len(d)


def grad_gather(planes, c, hw_out):
    px = c * 2
    return ops.gather_view(planes, (c, hw_out, 1, hw_out, N), (px, px, hw_out * px, hw_out * hw_out * px), hw_out, hw_out, N,
                           [(64 * j, 0, 0, 0) for j in range((c + 63) // 64)])


gen = torch.Generator(device="cuda").manual_seed(0)
P3 = ops.PAIRS[3]
for layer, hw_in, c_in, K, hw_out, C in (("conv2", 20, 32, 512, 9, 64), ("conv3", 9, 64, 576, 7, 32)):
    rows = N * hw_out * hw_out
    x = ops.split_planes(torch.rand((N * hw_in * hw_in, c_in), device="cuda", generator=gen), 2)
    pm = ops.split_planes(torch.rand((rows, K), device="cuda", generator=gen), 2)
    g = ops.split_planes(torch.randn((rows, C), device="cuda", generator=gen), 2)
    ga = ops.nature_cnn_gather(x, layer, N)
    gb = grad_gather(g, C, hw_out)
    print(layer, "SYRK   patch matrix %.1f us | gathered %.1f us" % (
        timed(lambda: ops.gemm(pm, pm, K, K, rows, trans=True, symmetric=True, pairs=P3)),
        timed(lambda: ops.gemm(x, x, K, K, rows, trans=True, symmetric=True, pairs=P3, a_gather=ga))), flush=True)
    print(layer, "wgrad  patch matrix %.1f us | gathered %.1f us" % (
        timed(lambda: ops.gemm(pm, g, K, C, rows, trans=True, pairs=P3)),
        timed(lambda: ops.gemm(x, g, K, C, rows, trans=True, pairs=P3, a_gather=ga, b_gather=gb))), flush=True)
    del pm
obs = torch.randint(0, 256, (N, 84, 84, 4), dtype=torch.uint8, device="cuda", generator=gen)
pc = ops.obs_pairs(obs)
rows = N * 400
p1 = [torch.randint(0, 256, (rows, 256), device="cuda", generator=gen).to(torch.bfloat16)]
g = ops.split_planes(torch.randn((rows, 32), device="cuda", generator=gen), 2)
ga = ops.nature_cnn_gather([pc], "conv1", N)
gb = grad_gather(g, 32, 20)
print("conv1 SYRK   patch matrix %.1f us | gathered %.1f us" % (
    timed(lambda: ops.gemm(p1, p1, 256, 256, rows, trans=True, symmetric=True, pairs=[(0, 0)])),
    timed(lambda: ops.gemm([pc], [pc], 256, 256, rows, trans=True, symmetric=True, pairs=[(0, 0)], a_gather=ga, perm_m=1, perm_n=1))))
print("conv1 wgrad  patch matrix %.1f us | gathered %.1f us" % (
    timed(lambda: ops.gemm(p1, g, 256, 32, rows, trans=True, pairs=[(0, 0), (0, 1)])),
    timed(lambda: ops.gemm([pc], g, 256, 32, rows, trans=True, pairs=[(0, 0), (0, 1)], a_gather=ga, b_gather=gb, perm_m=1))))
