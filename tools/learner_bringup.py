"""GPU bring-up of the learner engine: prints per-quantity errors against the fp64 oracle (no asserts) and a
first timing of the 32x20 ACKTR update.  Writes gpurun_out/learner_bringup.json."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import learner_checks as LC  # noqa: E402
import synth  # noqa: E402
from actorcritic_b200 import engine as eng  # noqa: E402

report = {}


def section(name, fn):
    t0 = time.time()
    try:
        report[name] = fn()
    except Exception:  # noqa: BLE001
        report[name] = {"error": traceback.format_exc()}
    report[name + "_sec"] = time.time() - t0
    print("==", name, json.dumps(report[name], indent=None, default=float)[:6000], flush=True)


def compute_parity(precision=0, e_count=4, t_count=5, c3=32, obs_kind="uniform"):
    cfg = eng.EngineConfig(num_envs=e_count, num_steps=t_count, conv3_filters=c3, precision=precision)
    e, o = LC.make_pair(cfg, seed=1)
    e.set_state(30, 0, False)
    o.global_step = 30
    n = e_count * t_count
    batch = synth.rollout(7, e_count, t_count, 4, obs_kind=obs_kind)
    y_hat, eps = synth.fisher_samples(9, n)
    e.load_batch(batch["observations"], batch["bootstrap_observations"], batch["actions"], batch["rewards"], batch["terminals"])
    e.phase1(torch.from_numpy(y_hat).cuda(), torch.from_numpy(eps).cuda())
    info = o.compute(batch, y_hat, eps, need_fisher=True)
    return LC.compare_compute(e, info, cfg, True)


def schedule(acktr=True):
    if acktr:
        cfg = eng.EngineConfig(num_envs=4, num_steps=5, conv3_filters=32, num_cold_updates=4, invert_every=2)
        return LC.run_schedule(cfg, 10)
    cfg = eng.EngineConfig.a2c(num_envs=4, num_steps=5)
    return LC.run_schedule(cfg, 3)


def timing(e_count=32, t_count=20, c3=32, precision=0, updates=30):
    cfg = eng.EngineConfig(num_envs=e_count, num_steps=t_count, conv3_filters=c3, precision=precision)
    e = eng.Engine(cfg)
    e.set_params(eng.orthogonal_init(4, c3, 0))
    e.set_state(40, 0, False)
    batch = synth.rollout(3, e_count, t_count, 4)
    e.load_batch(batch["observations"], batch["bootstrap_observations"], batch["actions"], batch["rewards"], batch["terminals"])
    for _ in range(12):
        e.update(fetch=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t1 = t2 = 0.0
    with torch.cuda.stream(e.stream):
        for _ in range(updates):
            ev[0].record()
            e.phase1()
            ev[1].record()
            e.phase2()
            ev[2].record()
            torch.cuda.synchronize()
            t1 += ev[0].elapsed_time(ev[1])
            t2 += ev[1].elapsed_time(ev[2])
    s = e.fetch_scalars()
    return dict(phase1_ms=t1 / updates, phase2_ms=t2 / updates, scalars=s, gs=e.global_step,
                launches=int(e.lib.acx_launch_count()))


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    print(torch.cuda.get_device_name(0), flush=True)
    for p in (0, 1, 2, 3):
        section("compute_p%d" % p, lambda: compute_parity(p))
    section("compute_p0_sparse", lambda: compute_parity(0, obs_kind="sparse"))
    section("schedule_acktr", lambda: schedule(True))
    section("schedule_a2c", lambda: schedule(False))
    for p in (0, 1, 2, 3):
        section("timing_32x20_p%d" % p, lambda: timing(precision=p))
    section("compute_p0_32x20", lambda: compute_parity(0, 32, 20))
    section("compute_p1_32x20", lambda: compute_parity(1, 32, 20))
    json.dump(report, open("gpurun_out/learner_bringup.json", "w"), indent=1, default=float)


if __name__ == "__main__":
    main()
