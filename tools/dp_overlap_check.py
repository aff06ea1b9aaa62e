"""torchrun --nproc-per-node 2 tools/dp_overlap_check.py: the overlapped all-reduce (input-factor prefix on a second stream and
communicator while phase 1 runs) must give bit-identical parameters to the single all-reduce after phase 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import synth  # noqa: E402
from actorcritic_b200 import engine as eng  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
envs, steps, updates = 8, 5, 26


def run(overlap, split="0"):
    os.environ["ACX_DP_SPLIT"] = split
    cfg = eng.EngineConfig(num_envs=envs, num_steps=steps, conv3_filters=32, world_size=world, num_cold_updates=4, invert_every=3,
                           seed=7)
    e = eng.Engine(cfg)
    e.set_params(eng.orthogonal_init(4, 32, seed=0))
    early = 0
    for u in range(updates):
        b = synth.rollout(1000 * rank + u, envs, steps, 4, obs_kind="sparse")
        e.load_batch(b["observations"], b["bootstrap_observations"], b["actions"], b["rewards"], b["terminals"])
        y, eps = synth.fisher_samples(50 + u + 100 * rank, envs * steps)
        e.phase1(torch.from_numpy(y).cuda(), torch.from_numpy(eps).cuda())
        e.allreduce(overlap=overlap)
        early += int(overlap and e.lib.acx_learner_wait_input_factors(e._h, None) == 0)
        e.phase2()
    torch.cuda.synchronize()
    sums = e.buffer("factor_sums", torch.float32).cpu().numpy().copy()
    return np.concatenate([e.get_params_flat().copy(), sums]), early


p_plain, _ = run(False)
p_over, early = run(True)
p_split, _ = run(False, split="1")
same = bool(np.array_equal(p_plain, p_over)) and bool(np.array_equal(p_plain, p_split))
t = torch.tensor([int(same)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("overlapped == plain == split exchange (parameters and factor sums):", bool(t.item()), "| split max |diff|",
          float(np.abs(p_plain - p_split).max()), "| overlapped:", "| updates with an early all-reduce:", early, "of", updates,
          "| max |diff|", float(np.abs(p_plain - p_over).max()))
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
