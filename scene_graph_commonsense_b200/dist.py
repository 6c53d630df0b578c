"""Multi-GPU plumbing: images are independent units, so ranks own whole images and the only exchange is ONE integer
all-reduce of the 765-slot counter vector per evaluation window.  The reference has no cross-rank metric
reduction at all (every rank writes its own results JSON, utils.py:486); SGB all_gathers pickled BoxLists
(SGB/maskrcnn_benchmark/utils/comm.py:48-91).  `torch.distributed` (NCCL over NVLink on GPUs, gloo in CPU tests) is
the transport; the counters never leave the device on the NCCL path.
"""
import os

import torch
import torch.distributed as dist


def shard_image_ids(image_ids, rank, world_size):
    """Round-robin deal identical to DistributedSampler(shuffle=False) (evaluate.py:44), without drop_last."""
    return list(image_ids)[rank::world_size]


def init_from_env(backend=None):
    """One process per GPU; reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* set by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_counters(counters):
    """In-place SUM of the int64 counter vector over all ranks (no-op for a single process)."""
    if counters.dtype != torch.int64:
        raise TypeError("counters must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
