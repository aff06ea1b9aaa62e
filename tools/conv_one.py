"""One acx_conv launch per case at the headline sizes (32 x 20: 672 forward samples, 1280 backward samples), for ncu.
usage: conv_one.py [f2|f3|d2|d3] [pairs] [reps] [planes]   (planes defaults to 2 for <= 3 pairs - the learner's default - else 3)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from actorcritic_b200 import ops  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "f2"
if case == "f1":   # conv1 forward on the row-pair copy of 672 observations (the update's forward batch)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    gen = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randint(0, 256, (672, 84, 84, 4), dtype=torch.uint8, device="cuda", generator=gen)
    pc = ops.obs_pairs(obs)
    w = torch.randn((256, 32), device="cuda", generator=gen) * 0.05
    bias = torch.zeros(32, device="cuda")
    from actorcritic_b200 import _lib
    import ctypes
    lib = _lib.load()
    f = torch.arange(256, device="cuda")
    col = ((f // 32) // 2) * 64 + ((f // 4) % 8) * 8 + ((f // 32) % 2) * 4 + f % 4
    wt = torch.empty((32, 256), device="cuda")
    wt[:, col] = w.t()
    wp = ops.split_planes(wt, 3)
    outs = [torch.zeros((672 * 400, 32), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    pa, pb = (ctypes.c_int * 2)(0, 0), (ctypes.c_int * 2)(0, 1)
    ws, os_ = ops._planes_struct(wp, 32, 256), ops._planes_struct(outs, 672 * 400, 32)
    run = lambda: _lib.check(lib.acx_conv1_pairs_forward(pc.data_ptr(), ctypes.byref(ws), 672, bias.data_ptr(), ctypes.c_float(1 / 255.0),
                                                         ctypes.byref(os_), 2, pa, pb, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        run()
    ev1.record()
    torch.cuda.synchronize()
    print(case, "us per call", 1e3 * ev0.elapsed_time(ev1) / reps, "dbg", os.environ.get("ACX_CONV_DEBUG", "0"))
    if int(os.environ.get("ACX_CONV_DEBUG", "0")) & 32:
        arr = (ctypes.c_longlong * 8)()
        lib.acx_debug_conv_trace(arr)
        t = list(arr)
        print("  CTA 0 MMA warp: total %d cycles over %d tiles; waiting for operands %d (%.0f%%), for a drained accumulator %d (%.0f%%); "
              "epilogue warp: waiting %d of %d cycles" % (t[0], t[3], t[1], 100.0 * t[1] / max(t[0], 1), t[2], 100.0 * t[2] / max(t[0], 1), t[5], t[7]))
    sys.exit(0)
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
planes = int(sys.argv[4]) if len(sys.argv) > 4 else (2 if npairs <= 3 else 3)
geom = {"f2": (20, 32, 4, 2, 9, 64), "d2": (20, 32, 4, 2, 9, 64), "f3": (9, 64, 3, 1, 7, 32), "d3": (9, 64, 3, 1, 7, 32)}[case]
hw_in, c_in, k, s, hw_out, c_out = geom
dgrad = case[0] == "d"
samples = 1280 if dgrad else 672
gen = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((k * k * c_in, c_out), device="cuda", generator=gen) * 0.05
pairs = ops.PAIRS[npairs]
if dgrad:
    x = torch.randn((samples * hw_out * hw_out, c_out), device="cuda", generator=gen)
    xp = [p.reshape(samples, hw_out, hw_out, c_out) for p in ops.split_planes(x, planes)]
    wp = ops.conv_dgrad_weights(w, geom)
    act = torch.rand((samples // 2, hw_in, hw_in, c_in), device="cuda", generator=gen).to(torch.bfloat16)
    outs = ops.conv(xp, wp, geom, samples, dgrad=True, mask=act, mask_samples=samples // 2, pairs=pairs, out_planes=planes)
    run = lambda: ops.conv(xp, wp, geom, samples, dgrad=True, mask=act, mask_samples=samples // 2, pairs=pairs, outs=outs)
else:
    x = torch.rand((samples * hw_in * hw_in, c_in), device="cuda", generator=gen)
    xp = [p.reshape(samples, hw_in, hw_in, c_in) for p in ops.split_planes(x, planes)]
    wp = ops.split_planes(w.t().contiguous(), planes)
    bias = torch.zeros(c_out, device="cuda")
    outs = ops.conv(xp, wp, geom, samples, bias=bias, relu=True, pairs=pairs, out_planes=planes)
    run = lambda: ops.conv(xp, wp, geom, samples, bias=bias, relu=True, pairs=pairs, outs=outs)
for _ in range(2):
    run()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    run()
ev1.record()
torch.cuda.synchronize()
print(case, "pairs", npairs, "dbg", os.environ.get("ACX_CONV_DEBUG", "0"), "us per call", 1e3 * ev0.elapsed_time(ev1) / reps)

if int(os.environ.get("ACX_CONV_DEBUG", "0")) & 32:
    import ctypes
    from actorcritic_b200 import _lib
    arr = (ctypes.c_longlong * 8)()
    _lib.load().acx_debug_conv_trace(arr)
    t = list(arr)
    print("  CTA 0 MMA warp: total %d cycles over %d tiles; waiting for operands %d (%.0f%%), for a drained accumulator %d (%.0f%%); "
          "epilogue warp: waiting %d of %d cycles" % (t[0], t[3], t[1], 100.0 * t[1] / max(t[0], 1), t[2], 100.0 * t[2] / max(t[0], 1),
                                                     t[5], t[7]))
