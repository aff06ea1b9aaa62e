"""Run K-PRE on E environments a few times (for ncu).  usage: pre_one.py [envs] [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from actorcritic_b200 import ops
pe = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ra = torch.randint(0, 256, (pe, 210, 160, 3), dtype=torch.uint8, device="cuda")
rb = torch.randint(0, 256, (pe, 210, 160, 3), dtype=torch.uint8, device="cuda")
stk = torch.randint(0, 256, (pe, 84, 84, 4), dtype=torch.uint8, device="cuda")
out = torch.empty_like(stk)
for _ in range(iters):
    ops.preprocess_stack(ra, rb, stk, out=out, out_env_stride=28224)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.preprocess_stack(ra, rb, stk, out=out, out_env_stride=28224)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("envs", pe, "ms", ms, "GB/s", pe * 258048 / ms / 1e6)
