"""Data-parallel plumbing: one process per GPU, rays sharded across ranks, replicated parameters, ONE
exchange per step (NCCL allreduce of the flat 2.19 MB gradient inside libnerfb200).  torch.distributed is used only for
rendezvous: shipping the NCCL unique id from rank 0 and for barriers / max-over-ranks timing in bench.py.
The reference has no multi-GPU code at all (single `cudaSetDevice(0)`, ANU/AcceleratedMipNeRF.cpp:10)."""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n: int, rank: int, world: int):
    """Contiguous, near-equal slice [lo, hi) of n rays for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(rays: dict, pixels, rank: int, world: int):
    n = pixels.shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in rays.items()}, pixels[lo:hi]


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Ship `nbytes` from rank `src` to every rank over the default process group (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(nbytes, dtype=torch.uint8, device=device or "cpu")
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def attach(model, device=None):
    """Attach `model` to the job's NCCL communicator: rank 0 creates the unique id, everyone joins."""
    import torch.distributed as dist

    import nerf_or_nothing_b200 as nb

    rank, world = dist.get_rank(), dist.get_world_size()
    uid = nb.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, nb.COMM_ID_BYTES, 0, device)
    model.comm_init(uid, rank, world)
    return rank, world


def global_loss_scale(local_loss_mult_sum: float) -> float:
    """Factor that turns a gradient normalised by the LOCAL sum(lossMult) into the global one (what the library does
    on the device with a 4-byte allreduce before the backward pass)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([local_loss_mult_sum], dtype=torch.float64)
    dist.all_reduce(t)
    return float(local_loss_mult_sum / t.item())


def allreduce_numpy(x: np.ndarray) -> np.ndarray:
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()
