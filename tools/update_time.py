"""python tools/update_time.py [steps]: ms per steady-state 32 x 20 ACKTR update (device-resident inputs, CUDA graphs) -
a light version of bench.py's `value` leg for knob sweeps (ACX_* environment variables)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import synth  # noqa: E402
from actorcritic_b200 import engine as eng  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
e = eng.Engine(eng.EngineConfig(num_envs=32, num_steps=20, conv3_filters=32, seed=1, num_lanes=int(os.environ.get('LANES', '0'))))
e.set_params(eng.orthogonal_init(4, 32, seed=0))
batches = [synth.rollout(10 + i, 32, 20, 4) for i in range(8)]
res = [{k: torch.from_numpy(v if v.dtype != bool else v.astype("uint8")).cuda() for k, v in b.items()} for b in batches]
e.set_state(30, 0, False)
with torch.cuda.stream(e.stream):
    for i in range(33):
        e.update(res[i % 8], fetch=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        e.update(res[i % 8], fetch=False)
    b.record()
    torch.cuda.synchronize()
knobs = {k: v for k, v in os.environ.items() if k.startswith("ACX_") or k == "LANES"}
print("ms_per_update %.4f" % (a.elapsed_time(b) / steps), knobs)
