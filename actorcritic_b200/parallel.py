"""Data-parallel plumbing: one process per GPU, environments sharded by rank, ONE collective per
update - a sum all-reduce over the engine's flat fp32 bucket [gradients | A statistics | G statistics | 3 loss
scalars + loss]; every rank then applies the identical phase-2 update, so parameters never diverge and nothing is
broadcast.  torch.distributed (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests) is the transport."""
import os

import torch
import torch.distributed as dist


def shard_range(num_envs, rank, world_size):
    """Rank r of k owns environments [r*E/k, (r+1)*E/k)  (SURVEY 8(e2)); E must divide evenly so that the mean of the
    per-rank means equals the global mean."""
    if num_envs % world_size != 0:
        raise ValueError("num_envs (%d) must be divisible by the number of ranks (%d)" % (num_envs, world_size))
    per = num_envs // world_size
    return rank * per, (rank + 1) * per


def shard_batch(batch, rank, world_size):
    """Slice the five train-step inputs (batch-major [environment, step, ...]) to this rank's environments."""
    lo, hi = shard_range(len(batch["observations"]), rank, world_size)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_mean_(bucket, group=None):
    """In-place mean over ranks of a flat bucket (sum all-reduce, then 1/k - the engine folds the 1/k into phase 2
    on the device; this helper is the host-visible form used by tests and tools)."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        bucket.mul_(1.0 / world)
    return bucket


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local
