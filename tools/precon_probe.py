"""Worst preconditioned-update error over an 8-update ACKTR schedule (8 x 10 batch) against the fp64 oracle, for the current
ACX_PRECON_LEVEL (plane-pair level of fc4's two preconditioning GEMMs; measured: level 1 = 3e-5 on fc4 instead of 1.7e-5,
and no faster - the GEMMs are not on the critical path - so the default stays at 2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import learner_checks as LC
from actorcritic_b200 import engine as eng
cfg = eng.EngineConfig(num_envs=8, num_steps=10, conv3_filters=32, num_cold_updates=2, invert_every=2)
recs = LC.run_schedule(cfg, 8, obs_kind="sparse")
worst = {}
for r in recs:
    for k, v in r.get("precon", {}).items():
        worst[k] = max(worst.get(k, 0), v)
print("level", os.environ.get("ACX_PRECON_LEVEL", "2"), "worst precon errors", {k: float("%.2g" % v) for k, v in worst.items()}, "step_rel", max(r["step_rel"] for r in recs))
