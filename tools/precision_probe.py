"""Error of one phase-1 pass against the fp64 oracle for engine variants.  usage: precision_probe.py E T kind key=value ..."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import learner_checks as LC  # noqa: E402
import synth  # noqa: E402
from actorcritic_b200 import engine as eng  # noqa: E402

e_count, t_count, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
kw = {}
for a in sys.argv[4:]:
    k, v = a.split("=")
    kw[k] = int(v)
cfg = eng.EngineConfig(num_envs=e_count, num_steps=t_count, conv3_filters=32, **kw)
e, o = LC.make_pair(cfg, seed=1)
e.set_state(30, 0, False)
o.global_step = 30
n = e_count * t_count
batch = synth.rollout(7, e_count, t_count, 4, obs_kind=kind)
y_hat, eps = synth.fisher_samples(9, n)
e.load_batch(batch["observations"], batch["bootstrap_observations"], batch["actions"], batch["rewards"], batch["terminals"])
e.phase1(torch.from_numpy(y_hat).cuda(), torch.from_numpy(eps).cuda())
info = o.compute(batch, y_hat, eps, need_fisher=True)
res = LC.compare_compute(e, info, cfg, True)
print(e_count, t_count, kind, kw, {k: float("%.2g" % v["rel"]) for k, v in res.items() if k[0] in "gG"})
masks = LC.engine_relu_masks(e)
print("mask disagreement (count, fraction, worst |pre| / rms):", LC.mask_disagreement(masks, info["fwd"]))
info2 = o.compute(batch, y_hat, eps, need_fisher=True, masks=masks)
res2 = LC.compare_compute(e, info2, cfg, True)
print("with the engine's masks:", {k: float("%.2g" % v["rel"]) for k, v in res2.items() if k[0] in "gG"})
