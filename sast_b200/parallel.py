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


def parse_cpulist(text: str) -> List[int]:
    """'0-15,32-47' -> [0..15, 32..47] (the sysfs cpulist format)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process to the CPUs that are NUMA-local to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE
    any pinned host buffer is allocated: page-locked memory is then placed on the GPU's own NUMA node, and with one rank
    per GPU the ranks' host->device copies stop sharing one socket's memory controllers.  Best effort: returns what it
    did; never raises."""
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        try:
            with open(f"{base}/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
        except OSError:
            pass
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), pci=bdf)
    except Exception as exc:      # noqa: BLE001 -- containers may hide sysfs; the run goes on unbound
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


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
