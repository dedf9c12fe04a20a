"""Multi-GPU host logic: records are independent units, so ranks take contiguous record ranges and the
only exchange is one all-reduce of the tallies vector.

On GPUs the reduction is `ntg_comm_allreduce_tallies` (ncclAllReduce, ncclUint64/ncclSum over NVLink); the
NCCL unique id is distributed with whatever process group the launcher provides (torch.distributed here).
With a gloo group (CPU tests) the same vector is reduced by torch.distributed itself.
"""
import numpy as np

from . import TALLY_FIELDS


def shard_records(n_records, world_size, rank):
    """-> (first_record, n_records_of_rank): contiguous, balanced, covering [0, n_records) exactly."""
    base, extra = divmod(n_records, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def tallies_to_i64(t):
    return np.array([t[f] for f in TALLY_FIELDS], dtype=np.uint64).view(np.int64)


def tallies_from_i64(a):
    return {f: int(v) for f, v in zip(TALLY_FIELDS, np.asarray(a, dtype=np.int64).view(np.uint64))}


def broadcast_bytes(data, src=0):
    """broadcast a small bytes object from `src` over the default torch.distributed group"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(len(data) if data is not None else 128, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def init_nccl_comm(ctx):
    """Create libntgpu's own NCCL communicator; the id travels over the launcher's process group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = ctx.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0)
    ctx.comm_init(world, rank, uid)


def allreduce_tallies(t, ctx=None):
    """Sum tallies over all ranks (wrapping u64).  ctx with an initialised communicator -> NCCL through the
    C ABI; otherwise torch.distributed on the default group (gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(t)
    if ctx is not None:
        return ctx.comm_allreduce_tallies(t)
    v = torch.from_numpy(tallies_to_i64(t).copy())
    dist.all_reduce(v, op=dist.ReduceOp.SUM)      # int64 addition wraps like u64
    return tallies_from_i64(v.numpy())
