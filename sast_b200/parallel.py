"""Multi-GPU plumbing for batch- / stream-sharded inference: one process per GPU, frames (or whole
event streams) partitioned across ranks, **no collective on the data path** -- windows never interact
across frames, and a stream's LSTM state stays on the rank that owns the stream.  The only
collectives are control-plane: a barrier around timed regions and a MAX/SUM of a few scalars.
(The reference is single-GPU for inference, validation.py:42, and DDP for training, train.py:91-98.)"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> 0, 1, 0)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str, device=None):
    """Join the process group if WORLD_SIZE > 1 (rendezvous on MASTER_ADDR/PORT, 127.0.0.1 on one node)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of n_items for `rank`; the first n_items % world ranks get one more."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def assign_streams(n_streams: int, world: int) -> List[List[int]]:
    """Event streams -> ranks, round-robin (stream s lives on rank s % world for its whole life, so its
    recurrent state never moves)."""
    return [list(range(r, n_streams, world)) for r in range(world)]


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_scalars(values: Sequence[float], op: str = "max", device="cpu") -> List[float]:
    """All-reduce a few scalars (MAX for timings -- the slowest rank defines the step; SUM for counts)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()


def throughput(frames_per_rank_per_step: int, steps: int, seconds_max_over_ranks: float, world: int) -> float:
    """Whole-job frames/s of a weak-scaled run: every rank processes its own frames, time = slowest rank."""
    return frames_per_rank_per_step * world * steps / seconds_max_over_ranks
