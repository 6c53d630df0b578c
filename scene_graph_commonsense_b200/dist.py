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


class CounterComm:
    """The counter all-reduce of the C ABI (`hc_counts_allreduce`, include/hiercom_b200.h): one NCCL communicator per process,
    created through `hc_nccl_comm_create` from an id that rank 0 draws and `torch.distributed` (any backend) broadcasts once.
    The reduction itself is a single ncclAllReduce(sum, int64) on the CURRENT CUDA stream - no torch collective on the data path."""

    def __init__(self, rank, world_size, device):
        import ctypes as C
        from . import _lib
        self.rank, self.world = rank, world_size
        lib = _lib.load()
        ident = (C.c_char * 128)()
        if rank == 0:
            _lib.check(lib.hc_nccl_unique_id(ident), "hc_nccl_unique_id")
        cuda_bcast = dist.get_backend() == "nccl"
        t = torch.tensor(list(ident.raw), dtype=torch.uint8, device=device if cuda_bcast else "cpu")
        dist.broadcast(t, src=0)
        ident = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().tolist()))
        comm = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.hc_nccl_comm_create(ident, world_size, rank, C.byref(comm)), "hc_nccl_comm_create")
        self._lib, self._comm = lib, comm

    def allreduce_(self, counters):
        """In-place sum over ranks, enqueued on the current stream."""
        from . import ops
        if counters.dtype != torch.int64 or not counters.is_cuda or not counters.is_contiguous():
            raise TypeError("counters must be a contiguous int64 CUDA tensor")
        return ops.counts_allreduce(counters, self._comm.value)

    def close(self):
        if self._comm:
            self._lib.hc_nccl_comm_destroy(self._comm)
            self._comm = None


_COUNTER_COMM = {}


def counter_comm(device):
    """The process's CounterComm (created on first use), or None for a single process / a CPU (gloo) run."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    device = torch.device(device)
    if device.type != "cuda" or os.environ.get("HC_COUNTS_ALLREDUCE", "abi") != "abi":
        return None
    key = (device.index, dist.get_world_size())
    if key not in _COUNTER_COMM:
        _COUNTER_COMM[key] = CounterComm(dist.get_rank(), dist.get_world_size(), device)
    return _COUNTER_COMM[key]


def allreduce_counters(counters, out=None):
    """SUM of the int64 counter vector over all ranks -> a NEW tensor (or `out`); the rank-local `counters` are left untouched.

    Counters are cumulative (the reference's Evaluator never resets them, evaluator.py:568-583), so reducing them IN PLACE more
    than once would multiply earlier totals by the world size; keeping the local vector intact makes the call idempotent: call it
    whenever global metrics are wanted (`pipeline.metrics_from_counters(allreduce_counters(pipe.counters))`).
    CUDA tensors go through the C ABI's `hc_counts_allreduce` (one ncclAllReduce on the current stream); CPU tensors (gloo tests)
    through torch.distributed.  A single process returns a copy."""
    if counters.dtype != torch.int64:
        raise TypeError("counters must be int64")
    if out is None:
        out = torch.empty_like(counters)
    out.copy_(counters)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        comm = counter_comm(out.device) if out.is_cuda else None
        if comm is not None:
            comm.allreduce_(out)
        else:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


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
