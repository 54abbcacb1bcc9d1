"""Data-parallel plumbing: one process per GPU, rays sharded across ranks, replicated parameters, ONE
exchange per step (NCCL allreduce of the flat 2.19 MB gradient + sum(lossMult) + loss numerators inside libnerfb200).  torch.distributed is used only for
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


def allreduce_step_buffer(grads_local_mean: np.ndarray, local_loss_mult_sum: float, local_losses):
    """Host restatement of the library's ONE collective per step (include/nerfb200.h, nerf_mipnerf_comm_init): every rank
    contributes [un-normalised gradient | sum(lossMult) | per-level loss numerators]; after a single sum-allreduce the
    global 1 / sum(lossMult) turns the first part into the global mean gradient and the tail into the global losses.
    `grads_local_mean` / `local_losses` are normalised by the LOCAL sum(lossMult) (what a single-process step returns)."""
    g = np.asarray(grads_local_mean, np.float64)
    losses = np.asarray(local_losses, np.float64).reshape(-1)
    buf = np.concatenate([g * local_loss_mult_sum, [local_loss_mult_sum], losses * local_loss_mult_sum])
    buf = allreduce_numpy(buf)
    n = g.shape[0]
    inv = 1.0 / buf[n]
    return buf[:n] * inv, buf[n + 1:] * inv


def allreduce_numpy(x: np.ndarray) -> np.ndarray:
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()
