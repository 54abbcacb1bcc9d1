"""-m gpu, needs >= 2 devices (skipped otherwise): the data-parallel step ON HARDWARE.

tests/test_dist_cpu.py checks the host logic with the oracle standing in for the device; here two real ranks (one process
per GPU, NCCL) run `nerf_mipnerf_train_step` on their halves of a batch and must reproduce the single-GPU step on the
whole batch: same loss, same (globally normalised) gradient up to summation order, and — what replicated training
relies on — bit-identical parameters on every rank after every step, with exactly ONE collective per step.

Also here: two handles on two devices inside ONE process (the per-device kernel attributes of ADVICE r1)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import nerf_or_nothing_b200 as nb
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
NET = dict(n_samples=64)
R_RANK, STEPS = 96, 3


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(step, n):
    from nerf_or_nothing_b200.scene import synthetic_rays

    rays, pix = synthetic_rays(n, width=100, height=100, seed=50 + step)
    rays["loss_mults"] = np.random.default_rng(step).uniform(0.5, 1.5, n).astype(np.float32)  # normaliser differs per rank
    return rays, pix


def _args(rays, pix, lo, hi):
    return [rays[k][lo:hi] for k in ("origins", "directions", "radii", "nears", "fars", "loss_mults")] + [pix[lo:hi]]


def _worker(rank, world, port, precision, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nerf_or_nothing_b200 import dist as nd

    m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R_RANK, precision=precision, device=rank, **NET))
    opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes(), device=rank)
    nd.attach(m, device="cuda")
    res = {}
    for step in range(STEPS):
        rays, pix = _batch(step, world * R_RANK)
        lo, hi = nd.shard_range(world * R_RANK, rank, world)
        before = m.launch_count()
        res[f"loss{step}"] = m.train_step(opt, *_args(rays, pix, lo, hi), 1e-3)
        res[f"launches{step}"] = m.launch_count() - before
        res[f"grads{step}"] = m.get_gradients()
        res[f"params{step}"] = m.get_params()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_gpu_step(precision, tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), precision, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # the same steps on ONE GPU over the whole batch; ray r of the global batch draws the same Philox sampling uniforms
    # in both set-ups (rank k's counters start at k * n_rays)
    single = nb.AcceleratedMipNeRF(nb.default_config(n_rays=world * R_RANK, precision=precision, **NET))
    opt = nb.AcceleratedAdamOptimizer(single.GetLayerSizes())
    for step in range(STEPS):
        rays, pix = _batch(step, world * R_RANK)
        loss = single.train_step(opt, *_args(rays, pix, 0, world * R_RANK), 1e-3)
        g, p = single.get_gradients(), single.get_params()
        for k in range(world):
            np.testing.assert_array_equal(r[k][f"params{step}"], r[0][f"params{step}"])  # ranks stay bit-identical
            np.testing.assert_array_equal(r[k][f"grads{step}"], r[0][f"grads{step}"])
            assert float(r[k][f"loss{step}"]) == float(r[0][f"loss{step}"])              # the GLOBAL loss on every rank
        e_g, e_l = rel_err(r[0][f"grads{step}"], g), abs(float(r[0][f"loss{step}"]) - loss) / loss
        e_p = float(np.abs(r[0][f"params{step}"] - p).max())
        print(f"{precision} step {step}: 2-rank vs 1-GPU gradient {e_g:.2e}, loss {e_l:.2e}, params max abs diff {e_p:.2e}")
        if step == 0:  # identical parameters going in: only summation order (and, in bf16, nothing else) differs
            assert e_g <= (2e-5 if precision == "fp32_tc" else 2e-3)
            assert e_l <= 1e-5
        assert e_p <= 2.5e-3  # Adam moves a parameter by <= lr per step whatever the gradient noise does to its sign
    assert int(r[0]["launches1"]) == int(r[1]["launches1"])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_handles_on_two_devices_in_one_process():
    """The dynamic-shared-memory opt-in of the tcgen05 kernels is per (kernel, device): a second handle on another GPU of
    the same process must launch them too, and give the same bits."""
    rays, pix = _batch(0, 64)
    out = []
    for device in (0, 1):
        for precision in ("fp32_tc", "bf16"):
            m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=64, precision=precision, device=device, **NET))
            opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes(), device=device)
            loss = m.train_step(opt, *_args(rays, pix, 0, 64), 1e-3)
            img = m.render(*_args(rays, pix, 0, 64)[:5])[0]
            out.append((device, precision, loss, m.get_params(), img))
    for a, b in ((out[0], out[2]), (out[1], out[3])):
        assert a[1] == b[1] and a[2] == b[2]
        np.testing.assert_array_equal(a[3], b[3])
        np.testing.assert_array_equal(a[4], b[4])
