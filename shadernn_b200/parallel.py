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
    """The single collective of the whole engine: rank `src`'s packed (BN-folded, split-bf16) weight arena overwrites
    every other rank's, in place, over NCCL. Returns the number of bytes broadcast."""
    ptr, nbytes = model.weight_arena()
    t = arena_as_tensor(ptr, nbytes, device)
    broadcast_buffer(t, src)
    # NCCL ran on torch's stream; the engine computes on its own non-blocking stream, which has no implicit ordering with it:
    # the arena must be complete before the first forward pass reads it
    if t.is_cuda:
        torch.cuda.synchronize(t.device)
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
