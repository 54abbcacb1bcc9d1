"""CPU, world_size = 2 over gloo: the data-parallel host logic (ray sharding, id broadcast, global normaliser,
gradient allreduce) with the CPU oracle standing in for the device step.  The property the multi-GPU path relies on:
sum over ranks of the UN-normalised per-shard gradients divided by the allreduced sum(lossMult) — one collective —
equals the single-process gradient (and loss) of the whole batch, so every rank applies the identical Adam step."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_or_nothing_b200 import dist as nd
    from nerf_or_nothing_b200.scene import synthetic_rays
    from oracle import oracle as orc

    orc.set_threads(2)
    cfg = orc.default_config(n_samples=16, net_depth=4, net_width=32, net_depth_condition=1, net_width_condition=16,
                             skip_layer=2, deg_point=6, deg_view=2)
    R, S = 22, 16  # not divisible by 2 ranks x anything nice: exercises the uneven shard
    rays, pix = synthetic_rays(R, width=100, height=100)
    rays["loss_mults"] = np.random.default_rng(0).uniform(0.5, 1.5, R).astype(np.float32)
    u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, R, S + 1) for lv in range(2)])
    params = orc.init_params(cfg, 7)

    # id plumbing: 128 opaque bytes from rank 0 reach everyone unchanged
    payload = bytes(range(128)) if rank == 0 else None
    got = nd.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))

    lo, hi = nd.shard_range(R, rank, world)
    srays, spix = nd.shard_batch(rays, pix, rank, world)
    assert spix.shape[0] == hi - lo
    o = orc.train_gradient(cfg, params, srays, spix, u[:, lo:hi], prec="f64")
    # the library's protocol: ONE allreduce of [un-normalised gradient | sum(lm) | loss numerators] (nerfb200.h, comm_init)
    g, losses = nd.allreduce_step_buffer(o["grads"], float(srays["loss_mults"].astype(np.float64).sum()), o["loss"])
    full = orc.train_gradient(cfg, params, rays, pix, u, prec="f64")
    err = np.abs(g - full["grads"]).max() / np.abs(full["grads"]).max()
    err = max(err, float(np.abs(losses - np.asarray(full["loss"], np.float64)).max() / np.abs(full["loss"]).max()))
    p1, _, _ = orc.adam_step(params, g, np.zeros_like(g), np.zeros_like(g), 1e-3, 1, 0, prec="f64")
    np.save(os.path.join(out_dir, f"p{rank}.npy"), p1)
    np.save(os.path.join(out_dir, f"e{rank}.npy"), np.array([err, lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from nerf_or_nothing_b200.dist import shard_range

    for n in (1, 7, 4096, 32768, 640000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_equals_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    e0, e1 = np.load(tmp_path / "e0.npy"), np.load(tmp_path / "e1.npy")
    assert e0[0] <= 1e-12 and e1[0] <= 1e-12  # fp64: only summation order differs
    assert (e0[1], e0[2], e1[1], e1[2]) == (0, 11, 11, 22)
    np.testing.assert_array_equal(np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy"))  # identical Adam step
