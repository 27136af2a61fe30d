"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the ONE collective this
workload has — a broadcast of the packed weight arena at init. The forward path has no exchange step:
images are independent units, each rank runs its contiguous slice of the batch and keeps its outputs.

The reference has no multi-device support at all; this module is new surface, kept deliberately thin.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_distributed(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total, world_size, rank):
    """Contiguous slice [start, start+count) of `total` units for `rank`; the first total % world ranks get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(total), int(world_size))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


class _DevicePointer:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


def arena_as_tensor(ptr, nbytes, device):
    return torch.as_tensor(_DevicePointer(ptr, nbytes), device=device)


def broadcast_buffer(t, src=0):
    """Broadcast tensor `t` in place from rank `src` (no-op for a single process). Works on CPU tensors under gloo,
    which is what the CPU test suite exercises."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def broadcast_model_weights(model, device, src=0):
    """The single collective of the whole engine: rank `src`'s packed (BN-folded, split-fp16) weight arena overwrites every other
    rank's, in place. On GPUs this is the library's own C++ path - snnb_nccl_comm_create + snnb_bcast_weights = one ncclBroadcast
    on the engine's stream (include/snnb.h); torch.distributed only carries the 128-byte NCCL id to the other ranks. Without a GPU
    (the gloo CPU tests) the same bytes travel through torch.distributed itself. Returns the number of bytes broadcast."""
    ptr, nbytes = model.weight_arena()
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return nbytes
    if str(device).startswith("cuda"):
        import ctypes as C

        from ._lib import check, lib
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = C.create_string_buffer(128)
        if rank == src:
            check(lib().snnb_nccl_unique_id(idbuf), "snnb_nccl_unique_id")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(device)
        dist.broadcast(t, src=src)
        idbuf = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        comm = C.c_void_p()
        check(lib().snnb_nccl_comm_create(model.ctx.h, rank, world, idbuf, C.byref(comm)), "snnb_nccl_comm_create")
        try:
            check(lib().snnb_bcast_weights(model.h, comm, src), "snnb_bcast_weights")
            model.ctx.sync()  # the arena is complete before the first forward pass reads it
        finally:
            lib().snnb_nccl_comm_destroy(comm)
        return nbytes
    t = arena_as_tensor(ptr, nbytes, device)
    broadcast_buffer(t, src)
    return nbytes


def max_over_ranks(value, device="cpu"):
    """MAX-reduce a python float over ranks (bench timing rule: report the slowest rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
