"""Multi-GPU data parallelism for the hot path: one process per GPU, rays / points sharded, parameters replicated,
one all-reduce of dL/dparams per step (the reference has no counterpart, its DDP support is
"discarded", nr3d_lib/config.py:74-75).

Rays (and every sample pack they own) are independent through march -> encode -> composite, so the data path needs no
collective; only the LoTD parameter gradient is summed across ranks.  The functions work with any torch.distributed
backend (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests).
"""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment.  Returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``n`` units (rays or points) owned by ``rank``; the first ``n % world``
    ranks get one extra unit so that every unit is owned exactly once."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, near: torch.Tensor, far: torch.Tensor, rank: int, world: int):
    """This rank's contiguous block of rays; each rank later builds its own pack_infos starting at 0."""
    b, e = shard_range(rays_o.shape[0], rank, world)
    return rays_o[b:e], rays_d[b:e], near[b:e], far[b:e], (b, e)


def allreduce_param_grads(grad: torch.Tensor, world: Optional[int] = None, async_op: bool = False):
    """Sum dL/dparams over all ranks in place (one all-reduce of n_params values per step).  No-op for one process."""
    if not dist.is_available() or not dist.is_initialized():
        return None
    if (world or dist.get_world_size()) == 1:
        return None
    return dist.all_reduce(grad, op=dist.ReduceOp.SUM, async_op=async_op)


def max_over_ranks(value: float, device=None) -> float:
    """max of a python float over all ranks (device-side timing is reported as the slowest rank's)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    """sum of a python float over all ranks (e.g. the number of samples every rank marched this step)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shutdown():
    """Tear the default process group down (silences NCCL's leak warning at interpreter exit)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
