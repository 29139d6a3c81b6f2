"""Multi-GPU plumbing: batch-index sharding + one tiny reduction.

IK queries and reward rows are independent, so rank r simply owns the contiguous range
``shard_range(n, rank, world)``; no data-path collective exists.  The only exchange is the
final ``all_reduce(SUM)`` of the 4 counters (IK: n, converged, success, sum(iterations);
reward: n, placed, gripped, threshold-adjacent) and ``all_reduce(MAX)`` of the elapsed time.
The payload is 32 bytes, i.e. latency bound on NVLink/NVSwitch: NCCL via torch.distributed
is the right tool and a fused compute+collective kernel would buy nothing.

Works with the ``nccl`` backend (one process per GPU) and with ``gloo`` (CPU tests).
"""

from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank ``rank``: sizes differ by at most one row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n < 0:
        raise ValueError("n must be >= 0")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun env when WORLD_SIZE > 1."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum the int64[4] counter vector over all ranks (returns a new tensor on the same device)."""
    if counters.dtype != torch.int64:
        raise ValueError("counters must be int64")
    out = counters.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def reduce_max(value: float, device=None) -> float:
    """Max of a scalar over all ranks (elapsed time: the job is as slow as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def summarize_ik(counters: torch.Tensor) -> dict:
    """Success rate / mean iterations from the (already reduced) IK counters."""
    n, conv, succ, iters = (int(x) for x in counters.tolist())
    return dict(n=n, converged=conv, success=succ, iterations=iters,
                success_rate=(succ / n if n else 0.0), mean_iterations=(iters / n if n else 0.0))
