"""Multi-GPU plumbing: the batch axis (hyper-parameter cell x SNR x channel) is embarrassingly
parallel, so each rank runs a contiguous slice of the flattened sample axis and a single all-gather of
the per-step NMSE log follows the last kernel.  No collective inside the loop.

One process per GPU (torchrun); NCCL over NVLink on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (no-op for a single process).
    Returns (rank, world_size, local_rank)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(lr)
            kw["device_id"] = torch.device("cuda", lr)
        dist.init_process_group(backend, **kw)
    return rank, ws, lr


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` samples owned by `rank` (ceil-sized, last ranks may be short)."""
    per = (total + world_size - 1) // world_size
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def gather_columns(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather a [steps, B_local] log whose columns are this rank's `shard_range(total, ...)` slice into
    the full [steps, total] log (every rank gets it).  The only collective of the path."""
    rank, ws = world()
    if ws == 1:
        return local
    per = (total + ws - 1) // ws
    steps = local.shape[0]
    pad = torch.zeros((steps, per), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat(bufs, dim=1)[:, :total]


def run_sharded(fn: Callable[[int, int], torch.Tensor], total: int) -> torch.Tensor:
    """fn(lo, hi) -> [steps, hi-lo] NMSE log of global samples lo..hi-1 computed on this rank; returns the
    full [steps, total] log on every rank."""
    rank, ws = world()
    lo, hi = shard_range(total, rank, ws)
    return gather_columns(fn(lo, hi), total)


def gather_rows(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather a [B_local, ...] real tensor whose rows are this rank's `shard_range(total, ...)` slice into the
    full [total, ...] tensor (every rank gets it)."""
    rank, ws = world()
    if ws == 1:
        return local
    per = (total + ws - 1) // ws
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat(bufs, dim=0)[:total]
