"""-m gpu: SURVEY §8(f) rows 2-4 — the device-resident BinDataset (SN/BinDataset.cs:10-52), on-device batch draw + gather,
the training step fed from it, image error / PSNR (SN/MipHelpers.cs:672), the learning-rate schedule
(SN/MipHelpers.cs:758-773) and checkpoint save / resume.  Integer and gather work is checked bit-exactly."""
import numpy as np
import pytest

import nerf_or_nothing_b200 as nb
from nerf_or_nothing_b200.scene import pack_records, synthetic_rays, unpack_records
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SMALL = dict(n_samples=32, net_depth=4, net_width=64, net_depth_condition=1, net_width_condition=64, skip_layer=2,
             deg_point=8, deg_view=2)


def _records(n=5000, seed=3):
    rays, pix = synthetic_rays(n, width=100, height=100, n_views=4, seed=seed)
    rays["loss_mults"] = (1.0 + 0.25 * (np.arange(n) % 3)).astype(np.float32)  # exercise the lossmult column
    return pack_records(rays, pix)


def _oracle_indices(seed, step, first_slot, n_rays, n_records):
    """Philox4x32-10 restatement of the sampler: counter (slot, 0, step, 0x0DA7A5E7), key = seed, idx = floor(w0 * n / 2^32)."""
    out = np.empty(n_rays, np.int64)
    for i in range(n_rays):
        w = orc.philox([(first_slot + i) & 0xFFFFFFFF, 0, step, 0x0DA7A5E7], [seed & 0xFFFFFFFF, seed >> 32])
        out[i] = (int(w[0]) * n_records) >> 32
    return out


def test_draw_indices_match_philox_restatement_and_cover_the_range():
    rec = _records()
    ds = nb.BinDataset(rec)
    assert len(ds) == rec.shape[0]
    for seed, step, slot0 in ((2024, 0, 0), (2024, 7, 0), ((5 << 32) | 11, 3, 4096)):
        idx = ds.draw_indices(seed, step, 257, first_slot=slot0)
        np.testing.assert_array_equal(idx, _oracle_indices(seed, step, slot0, 257, len(ds)))
        assert idx.min() >= 0 and idx.max() < len(ds)
    a, b = ds.draw_indices(2024, 1, 4096), ds.draw_indices(2024, 2, 4096)
    assert (a != b).mean() > 0.99                       # a new batch every step
    assert len(np.unique(a)) > 0.6 * len(ds) * (1 - np.exp(-4096 / len(ds)))  # with replacement, spread over the file
    hist = np.bincount(ds.draw_indices(1, 0, 200000) * 10 // len(ds), minlength=10)
    assert np.all(np.abs(hist / 20000 - 1) < 0.05)      # uniform over the records


def test_gather_is_bit_exact_for_explicit_and_drawn_indices(tmp_path):
    rec = _records()
    path = tmp_path / "train_data.bin"
    rec.tofile(path)                                    # the reference's file format: raw 64-byte records
    ds = nb.BinDataset(path)
    assert len(ds) == rec.shape[0]
    idx = np.array([0, len(ds) - 1, 17, 17, 4242, 1], np.int64)  # duplicates and both ends
    got = ds.gather(len(idx), indices=idx)
    rays, pix = unpack_records(rec[idx])
    for k in ("origins", "directions", "radii", "nears", "fars", "loss_mults"):
        np.testing.assert_array_equal(got[k], rays[k])
    np.testing.assert_array_equal(got["pixels"], pix)
    drawn = ds.draw_indices(99, 5, 1000, first_slot=3)
    got = ds.gather(1000, seed=99, step=5, first_slot=3)
    rays, pix = unpack_records(rec[drawn])
    np.testing.assert_array_equal(got["origins"], rays["origins"])
    np.testing.assert_array_equal(got["pixels"], pix)
    with pytest.raises(nb.NerfError):
        nb.BinDataset(tmp_path / "missing.bin")
    (tmp_path / "short.bin").write_bytes(b"x" * 63)      # not even one record (SN/BinDataset.cs:38-39)
    with pytest.raises(nb.NerfError):
        nb.BinDataset(tmp_path / "short.bin")


def test_train_step_from_resident_dataset_equals_host_fed_step():
    R = 96
    rec = _records(3000)
    ds = nb.BinDataset(rec)
    cfg = nb.default_config(n_rays=R, precision="fp32", **SMALL)
    a, b = nb.AcceleratedMipNeRF(cfg), nb.AcceleratedMipNeRF(cfg)
    oa, ob = nb.AcceleratedAdamOptimizer(a.GetLayerSizes()), nb.AcceleratedAdamOptimizer(b.GetLayerSizes())
    for step in range(4):
        lr = nb.learning_rate_decay(step, lr_delay_steps=3)
        idx = ds.draw_indices(2024, step, R)
        rays, pix = unpack_records(rec[idx])
        la = a.train_step_dataset(oa, ds, R, 2024, lr)
        lb = b.train_step(ob, rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"], pix, lr)
        assert la == lb, (step, la, lb)
    np.testing.assert_array_equal(a.get_params(), b.get_params())


def test_checkpoint_resume_is_bitwise_identical(tmp_path):
    R = 64
    rec = _records(2000)
    ds = nb.BinDataset(rec)
    cfg = nb.default_config(n_rays=R, precision="fp32", **SMALL)
    m = nb.AcceleratedMipNeRF(cfg)
    opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
    for step in range(3):
        m.train_step_dataset(opt, ds, R, 7, 1e-3)
    ck = tmp_path / "step3.ckpt"
    m.save_checkpoint(ck, opt)
    tail_a = [m.train_step_dataset(opt, ds, R, 7, 1e-3) for _ in range(3)]
    m2 = nb.AcceleratedMipNeRF(cfg)
    opt2 = nb.AcceleratedAdamOptimizer(m2.GetLayerSizes())
    m2.set_params(np.zeros_like(m.get_params()))        # nothing of the fresh model may survive the load
    m2.load_checkpoint(ck, opt2)
    tail_b = [m2.train_step_dataset(opt2, ds, R, 7, 1e-3) for _ in range(3)]
    assert tail_a == tail_b                              # same batches (step counter restored), same Adam state
    np.testing.assert_array_equal(m.get_params(), m2.get_params())
    # parameters-only checkpoint; an optimizer cannot be restored from it
    m.save_checkpoint(tmp_path / "params.ckpt")
    m3 = nb.AcceleratedMipNeRF(cfg)
    m3.load_checkpoint(tmp_path / "params.ckpt")
    np.testing.assert_array_equal(m3.get_params(), m.get_params())
    with pytest.raises(nb.NerfError):
        m3.load_checkpoint(tmp_path / "params.ckpt", nb.AcceleratedAdamOptimizer(m3.GetLayerSizes()))
    # another network shape, a truncated file and a foreign file are rejected
    other = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, precision="fp32", **{**SMALL, "net_width": 128}))
    with pytest.raises(nb.NerfError):
        other.load_checkpoint(ck)
    data = ck.read_bytes()
    (tmp_path / "cut.ckpt").write_bytes(data[: len(data) // 2])
    with pytest.raises(nb.NerfError):
        m2.load_checkpoint(tmp_path / "cut.ckpt", opt2)
    (tmp_path / "junk.ckpt").write_bytes(b"\0" * 4096)
    with pytest.raises(nb.NerfError):
        m2.load_checkpoint(tmp_path / "junk.ckpt")


def test_generate_rays_and_render_view_match_oracle():
    """Dataset.GenerateRays on the device (SN/Dataset.cs:111-176): the same bits as the CPU restatement for both edge modes and
    a pixel sub-range; nerf_mipnerf_render_view (rays generated chunk by chunk on the device, only the pose crosses PCIe) gives
    the same image bits as rendering the oracle's host ray arrays."""
    from tests.test_oracle_cpu import _pose

    W, H = 53, 41
    c2w = _pose(3)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    for mode in (0, 1):
        got = nb.generate_rays(c2w, focal, W, H, edge_mode=mode)
        ref = orc.generate_rays(c2w, focal, W, H, edge_mode=mode)
        for k in ref:
            np.testing.assert_array_equal(got[k], ref[k])
    part = nb.generate_rays(c2w, focal, W, H, first_pixel=500, n_pixels=777)
    for k in ref:
        np.testing.assert_array_equal(part[k], ref[k][500:1277])
    for precision in ("fp32", "bf16"):
        m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=512, n_samples=32, precision=precision))  # 2173 pixels: 5 chunks, ragged tail
        rgb, depth, acc = m.render_view(c2w, focal, W, H)
        rgb2, depth2, acc2 = m.render(ref["origins"], ref["directions"], ref["radii"], ref["nears"], ref["fars"])
        np.testing.assert_array_equal(rgb, rgb2)
        np.testing.assert_array_equal(depth, depth2)
        np.testing.assert_array_equal(acc, acc2)
        assert rgb.shape == (W * H, 3) and np.isfinite(rgb).all()


def test_ssim_matches_oracle():
    """nerf_image_ssim (SN/MipHelpers.cs:688-737) against the CPU restatement: same taps in the same order in fp32, so the
    map agrees to rounding noise; sizes that are not multiples of the 16 x 16 tile, a different window, and the mean of an
    800 x 800 image (the render size of configs[3])."""
    rng = np.random.default_rng(3)
    for (H, W, fs, sigma) in ((37, 53, 11, 1.5), (64, 48, 7, 1.0), (800, 800, 11, 1.5)):
        a = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
        b = np.clip(a + rng.normal(0, 0.05, a.shape), 0, 1).astype(np.float32)
        mean, m = nb.image_ssim(a, b, filter_size=fs, filter_sigma=sigma, want_map=True)
        omean, om = orc.ssim(a, b, filter_size=fs, filter_sigma=sigma)
        print(f"ssim {H}x{W} window {fs}: {mean:.6f} (oracle {omean:.6f}), map max abs diff {np.abs(m - om).max():.2e}")
        np.testing.assert_allclose(m, om, atol=2e-6)
        assert abs(mean - omean) <= 1e-7
    assert abs(nb.image_ssim(a, a) - 1.0) <= 1e-5
    assert nb.image_ssim(a, b) == nb.image_ssim(a, b)  # block partials are added in block order


def test_image_error_and_psnr():
    rng = np.random.default_rng(0)
    a = rng.random((100, 120, 3), dtype=np.float32)
    b = np.clip(a + rng.normal(scale=0.05, size=a.shape).astype(np.float32), 0, 1)
    mse, psnr = nb.image_error(a, b)
    ref = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    assert abs(mse - ref) <= 1e-12 * ref
    assert abs(psnr - (-10.0 / np.log(10.0) * np.log(ref))) <= 1e-9    # MseToPsnr, SN/MipHelpers.cs:672
    assert nb.image_error(a, a)[0] == 0.0


def test_learning_rate_schedule_matches_oracle():
    for step in (0, 1, 100, 2499, 2500, 2501, 50000, 999999, 1000000, 2000000):
        ref = orc.learning_rate_decay(step)
        assert abs(nb.learning_rate_decay(step) - ref) <= 2e-6 * ref
    assert abs(nb.learning_rate_decay(10, 1e-2, 1e-4, 100, 0, 1.0) - orc.learning_rate_decay(10, 1e-2, 1e-4, 100, 0, 1.0)) <= 1e-8
