"""Multi-GPU plumbing for the rollout path: the env batch is sharded contiguously, one process per
GPU, with NO data-path collective (envs are independent).  torch.distributed is used only to agree
on timings / statistics (and, in the PPO update, for the gradient all-reduce)."""
import os

import torch
import torch.distributed as dist


def env_shard(total_envs, rank, world):
    """Contiguous shard [base, base+count) of `total_envs` global env ids for `rank` of `world`.
    The first (total % world) ranks take one extra env."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(int(total_envs), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def world_info():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init(backend=None):
    rank, local_rank, world = world_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, **kw)
    return rank, local_rank, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def reduce_scalar(value, op="max", device=None):
    """max / sum / min of a Python float over all ranks (identity when not distributed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())
