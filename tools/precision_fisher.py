"""Error of one phase-1 pass at 32 x 20 against the fp64 oracle (norm-relative), for the current ACX_FISHER_LEVEL.
usage: [ACX_FISHER_LEVEL=2] python tools/precision_fisher.py [obs_kind]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import learner_bringup as LB  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
res = LB.compute_parity(0, 32, 20, obs_kind=kind)
print("fisher_level", os.environ.get("ACX_FISHER_LEVEL", "default(1)"), kind, json.dumps(res, default=float))
