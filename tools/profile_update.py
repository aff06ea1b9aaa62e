"""Run a few steady-state ACKTR updates (32 envs x 20 steps) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  --updates N (default 1); --with-refresh makes the profiled window start at an
update that refreshes the inverses."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import synth  # noqa: E402
from actorcritic_b200 import engine as eng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--updates", type=int, default=1)
ap.add_argument("--envs", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--conv3", type=int, default=32)
ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--with-refresh", action="store_true")
ap.add_argument("--conv-impl", type=int, default=0)
ap.add_argument("--lanes", type=int, default=0)
args = ap.parse_args()

cfg = eng.EngineConfig(num_envs=args.envs, num_steps=args.steps, conv3_filters=args.conv3, precision=args.precision,
                       conv_impl=args.conv_impl, num_lanes=args.lanes)
e = eng.Engine(cfg)
e.set_params(eng.orthogonal_init(4, args.conv3, 0))
b = synth.rollout(3, args.envs, args.steps, 4)
e.load_batch(b["observations"], b["bootstrap_observations"], b["actions"], b["rewards"], b["terminals"])
e.set_state(30, 0, False)
for _ in range(11):
    e.update(fetch=False)
# gs is 41 now; updates that start at gs = 40 + 10k refresh the inverses
while (e.global_step - 30) % 10 != (0 if args.with_refresh else 1):
    e.update(fetch=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.updates):
    e.update(fetch=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.updates, "updates ending at gs", e.global_step, "launches", e.lib.acx_launch_count())
