"""World-size-N check of the peer-memory exchange (csrc/peer.cu; N > 2 takes the two-shot route): the same 7 updates with the
exchange over peer memory and with NCCL; every rank must hold bit-identical parameters in either mode, and the two modes must
agree to fp32 summation order.  Launch: torchrun --nproc-per-node N tools/peer_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from oracle import network as onet  # noqa: E402
from actorcritic_b200 import _lib, engine as eng, parallel  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
envs = 2 * world
res = {}
for peer in ("1", "0"):
    os.environ["ACX_PEER"] = peer
    e = eng.Engine(eng.EngineConfig(num_envs=envs // world, world_size=world, num_steps=5, num_cold_updates=2, invert_every=2,
                                    lr_decay_steps=1000.0))
    e.set_params(onet.perturbed_params(4, 32, 7))
    for u in range(7):
        batch = parallel.shard_batch(synth.rollout(60 + u, envs, 5, 4, obs_kind="sparse"), rank, world)
        y, eps = synth.fisher_samples(70 + u, envs * 5)
        lo, hi = parallel.shard_range(envs, rank, world)
        yy = torch.from_numpy(y.reshape(envs, 5)[lo:hi].reshape(-1).copy()).cuda()
        ee = torch.from_numpy(eps.reshape(envs, 5)[lo:hi].reshape(-1).copy()).cuda()
        e.update(batch, yy, ee, fetch=False)
    torch.cuda.synchronize()
    assert bool(getattr(e, "_peer_state", False)) == (peer == "1")
    assert _lib.load().acx_peer_error() == 0
    p = torch.from_numpy(e.get_params_flat().copy()).cuda()
    s = e.buffer("factor_sums").clone()
    allp = [torch.empty_like(p) for _ in range(world)]
    alls = [torch.empty_like(s) for _ in range(world)]
    dist.all_gather(allp, p)
    dist.all_gather(alls, s)
    for k in range(world):
        assert torch.equal(allp[k], allp[0]) and torch.equal(alls[k], alls[0]), "ranks diverged (peer=%s, rank %d)" % (peer, k)
    res[peer] = (p.cpu().numpy(), s.cpu().numpy())
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))  # noqa: E731
ep, es = rel(res["1"][0], res["0"][0]), rel(res["1"][1], res["0"][1])
if rank == 0:
    print("peer_check world=%d: ranks identical in both modes; peer vs NCCL: params %.2e, factor sums %.2e" % (world, ep, es))
assert ep <= 1e-4 and es <= 1e-5, (ep, es)
dist.destroy_process_group()
