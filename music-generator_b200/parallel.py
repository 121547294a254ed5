"""Data-parallel plumbing (one process per GPU): rendezvous, batch sharding and the
single flat-buffer gradient all-reduce.  The reference has no distributed code;
this is the multi-GPU row of SURVEY.md 8e.  Backend `nccl` on GPUs (NVLink 5 /
NVSwitch), `gloo` in the CPU tests."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Strided shard of a dataset of n sequences (every sequence lands on exactly one rank)."""
    return np.arange(rank, n, world)


def allreduce_flat(flat: torch.Tensor) -> torch.Tensor:
    """Sum the flat gradient buffer over ranks in place (one collective per step).
    The 1/world average is applied by the Nadam kernel (gscale)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
