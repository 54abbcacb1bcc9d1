"""CPU tests that PIN the oracle (the reference ships no tests/golden vectors — SURVEY §4):
P2 fp32 build vs fp64 shadow, P3 torch-fp64 autograd, P4 analytic identities + Philox KATs."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import torch_spec

SMALL = dict(n_samples=16, net_depth=4, net_width=32, net_depth_condition=1, net_width_condition=16,
             skip_layer=2, deg_point=6, deg_view=2)


def small_cfg(**kw):
    d = dict(SMALL)
    d.update(kw)
    return orc.default_config(**d)


def test_layer_table_matches_reference():
    # 544 768 weights + 2 180 biases = 546 948 (ANU/AcceleratedMLP.cpp:131-154)
    cfg = orc.default_config()
    sizes = orc.layer_sizes(cfg)
    assert len(sizes) == 22
    assert sizes[:11] == [256 * 96] + [256 * 256] * 3 + [256 * 352] + [256 * 256] * 3 + [256, 128 * 283, 3 * 128]
    assert sizes[11:] == [256] * 8 + [1, 128, 3]
    assert orc.num_params(cfg) == 546948


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = orc.sampling_uniforms(99, 3, 1, 0, 64, 17)
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.05


def test_sample_t_vals_stratified():
    R, S = 8, 16
    near, far = np.full(R, 2.0), np.full(R, 6.0)
    u = orc.sampling_uniforms(1, 0, 0, 0, R, S + 1)
    t = orc.sample_t_vals(near, far, u, S)
    s = 2.0 + 4.0 * np.arange(S + 1) / S
    mids = 0.5 * (s[1:] + s[:-1])
    lower, upper = np.concatenate([s[:1], mids]), np.concatenate([mids, s[-1:]])
    assert np.all(t >= lower - 1e-6) and np.all(t <= upper + 1e-6)
    assert np.all(np.diff(t, axis=1) >= 0)
    np.testing.assert_allclose(t, lower + u * (upper - lower), rtol=1e-6)
    t_det = orc.sample_t_vals(near, far, None, S, randomized=0)
    np.testing.assert_allclose(t_det, np.broadcast_to(s, (R, S + 1)), rtol=1e-6)


def test_resample_follows_pdf():
    R, S = 4, 32
    t = np.sort(np.random.default_rng(0).uniform(2, 6, (R, S + 1)), 1)
    w = np.zeros((R, S))
    w[:, 10] = 1.0  # all mass in bin 10 -> samples concentrate around it
    u = orc.sampling_uniforms(5, 0, 1, 0, R, S + 1)
    for prec in ("f32", "f64"):
        tn = orc.resample_t_vals(t, w, u, prec=prec)
        assert np.all(np.diff(tn, axis=1) >= 0), "sorted"
        assert np.all(tn >= t[:, :1] - 1e-6) and np.all(tn <= t[:, -1:] + 1e-6)
        inside = (tn >= t[:, 8:9]) & (tn <= t[:, 13:14])
        assert inside.mean() > 0.6  # blur spreads 1.0 over bins 9..11 (+0.01 padding everywhere)
    a, b = orc.resample_t_vals(t, w, u, prec="f32"), orc.resample_t_vals(t, w, u, prec="f64")
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5)
    # zero weights everywhere: padding only -> uniform pdf -> roughly the input partition
    tz = orc.resample_t_vals(t, np.zeros((R, S)), u, prec="f64")
    assert np.all(np.diff(tz, axis=1) >= 0)


def test_cast_rays_and_ipe_against_torch_fp64():
    rng = np.random.default_rng(1)
    R, S = 6, 16
    t = np.sort(rng.uniform(2, 6, (R, S + 1)), 1)
    o, d, rad = rng.normal(size=(R, 3)), rng.normal(size=(R, 3)), rng.uniform(1e-4, 1e-2, R)
    mean64, cov64 = orc.cast_rays(t, o, d, rad, prec="f64")
    tm, tc = torch_spec.cast_rays(*[torch.tensor(x) for x in (t, o, d, rad)])
    np.testing.assert_allclose(mean64, tm.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(cov64, tc.numpy(), rtol=1e-9, atol=1e-15)
    mean32, cov32 = orc.cast_rays(t, o, d, rad, prec="f32")
    np.testing.assert_allclose(mean32, mean64, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cov32, cov64, rtol=2e-3, atol=1e-9)
    enc64 = orc.encode_position(mean64, cov64, 16, prec="f64")
    te = torch_spec.ipe(torch.tensor(mean64), torch.tensor(cov64), 16).reshape(-1, 96)
    np.testing.assert_allclose(enc64, te.numpy(), rtol=1e-9, atol=1e-12)
    # layout [f][sin xyz, cos xyz] (.cu:199-204)
    m0, c0 = mean64.reshape(-1, 3)[0], cov64.reshape(-1, 3)[0]
    assert np.isclose(enc64[0, 2 * 6 + 1], np.exp(-0.5 * c0[1] * 16) * np.sin(m0[1] * 4))
    assert np.isclose(enc64[0, 2 * 6 + 3 + 2], np.exp(-0.5 * c0[2] * 16) * np.cos(m0[2] * 4))
    # identity: var -> 0 gives plain sin/cos
    enc0 = orc.encode_position(mean64, np.zeros_like(cov64), 4, prec="f64")
    np.testing.assert_allclose(enc0[:, :3], np.sin(mean64.reshape(-1, 3)), atol=1e-12)
    ed = orc.encode_direction(d, 4, prec="f64")
    np.testing.assert_allclose(ed, torch_spec.dir_enc(torch.tensor(d), 4).numpy(), atol=1e-12)
    assert ed.shape == (R, 27)


def test_rendering_identities_and_gradient():
    rng = np.random.default_rng(2)
    R, S = 5, 24
    rgb, den = rng.uniform(0, 1, (R, S, 3)), rng.uniform(0, 3, (R, S))
    t = np.sort(rng.uniform(2, 6, (R, S + 1)), 1)
    d = rng.normal(size=(R, 3))
    o = orc.volumetric_rendering(rgb, den, t, d, prec="f64")
    # sum(w) + T_end = 1
    T_end = o["transmittance"][:, -1] * (1 - o["alpha"][:, -1])
    np.testing.assert_allclose(o["acc"] + T_end, 1.0, atol=1e-12)
    comp, acc, w = torch_spec.render(*[torch.tensor(x) for x in (rgb, den, t, d)])
    np.testing.assert_allclose(o["comp_rgb"], comp.numpy(), atol=1e-12)
    np.testing.assert_allclose(o["weights"], w.numpy(), atol=1e-12)
    assert np.all(o["depth"] >= t[:, 0]) and np.all(o["depth"] <= t[:, -1])
    o32 = orc.volumetric_rendering(rgb, den, t, d, prec="f32")
    np.testing.assert_allclose(o32["comp_rgb"], o["comp_rgb"], rtol=1e-5, atol=1e-6)
    # gradient vs autograd (exact mode)
    g = rng.normal(size=(R, 3))
    trgb, tden = torch.tensor(rgb, requires_grad=True), torch.tensor(den, requires_grad=True)
    comp, _, _ = torch_spec.render(trgb, tden, torch.tensor(t), torch.tensor(d))
    (comp * torch.tensor(g)).sum().backward()
    d_rgb, d_den = orc.volumetric_rendering_gradient(g, rgb, den, t, d, prec="f64")
    np.testing.assert_allclose(d_rgb, trgb.grad.numpy(), atol=1e-12)
    np.testing.assert_allclose(d_den, tden.grad.numpy(), atol=1e-11)
    # reference mode (A-D12): last sample gets nothing, its term is dropped from the others
    r_rgb, r_den = orc.volumetric_rendering_gradient(g, rgb, den, t, d, last_sample_mode=1, prec="f64")
    assert np.all(r_rgb[:, -1] == 0) and np.all(r_den[:, -1] == 0)
    np.testing.assert_allclose(r_rgb[:, :-1], d_rgb[:, :-1], atol=1e-12)
    assert not np.allclose(r_den[:, :-1], d_den[:, :-1])


def test_output_gradient_and_activations():
    rng = np.random.default_rng(3)
    R = 7
    comp, pix, lm = rng.uniform(0, 1, (R, 3)), rng.uniform(0, 1, (R, 3)), rng.uniform(0.5, 2, R)
    g = orc.output_gradient(comp, pix, lm, lm.sum(), 0.1, prec="f64")
    np.testing.assert_allclose(g, 2 * 0.1 * lm[:, None] / lm.sum() * (comp - pix), rtol=1e-12)
    cfg = orc.default_config(density_bias=-1.0, rgb_padding=0.001)
    rd, rr = rng.normal(size=20), rng.normal(size=(20, 3))
    den, rgb = orc.output_activations(cfg, rd, rr, prec="f64")
    np.testing.assert_allclose(den, np.log1p(np.exp(rd - 1)), rtol=1e-12)
    np.testing.assert_allclose(rgb, 1 / (1 + np.exp(-rr)) * 1.002 - 0.001, rtol=1e-12)
    dd, dr = orc.output_activations_grad(cfg, rd, rr, np.ones(20), np.ones((20, 3)), prec="f64")
    np.testing.assert_allclose(dd, 1 / (1 + np.exp(-(rd - 1))), rtol=1e-12)
    s = 1 / (1 + np.exp(-rr))
    np.testing.assert_allclose(dr, s * (1 - s) * 1.002, rtol=1e-12)


def test_adam_closed_form_step1():
    rng = np.random.default_rng(4)
    n = 1000
    p, g = rng.normal(size=n), rng.normal(size=n) * 1e-2
    p1, m1, v1 = orc.adam_step(p, g, np.zeros(n), np.zeros(n), 1e-3, 1, eps_mode=0, prec="f64")
    # dp = -lr * g / sqrt(g^2 + 1e-8)
    np.testing.assert_allclose(p1 - p, -1e-3 * g / np.sqrt(g * g + 1e-8), rtol=1e-9)
    np.testing.assert_allclose(m1, 0.1 * g, rtol=1e-12)
    p1b, _, _ = orc.adam_step(p, g, np.zeros(n), np.zeros(n), 1e-3, 1, eps_mode=1, prec="f64")
    np.testing.assert_allclose(p1b - p, -1e-3 * g / (np.abs(g) + 1e-8), rtol=1e-9)
    p32, _, _ = orc.adam_step(p, g, np.zeros(n), np.zeros(n), 1e-3, 1, prec="f32")
    np.testing.assert_allclose(p32, p1, rtol=1e-5, atol=1e-6)


def test_lr_schedule():
    # SN/MipHelpers.cs:758-773 with SN/TrainState.cs:54-57 defaults
    assert np.isclose(orc.learning_rate_decay(0), 0.01 * 5e-4, rtol=1e-5)
    assert np.isclose(orc.learning_rate_decay(2500), 5e-4 * np.exp(np.log(5e-6 / 5e-4) * 2500 / 1e6), rtol=1e-5)
    assert np.isclose(orc.learning_rate_decay(1000000), 5e-6, rtol=1e-4)


def test_mlp_forward_backward_vs_autograd():
    cfg = small_cfg()
    shapes = orc.layer_shapes(cfg)
    rng = np.random.default_rng(5)
    P, Dd, M = 6 * cfg.deg_point, 3 + 6 * cfg.deg_view, 50
    params = orc.init_params(cfg, 7).astype(np.float64)
    params[-sum(shapes[0]):] = rng.normal(size=sum(shapes[0])) * 0.1  # non-zero biases
    ep, ed = rng.normal(size=(M, P)), rng.normal(size=(M, Dd))
    rd, rr, acts = orc.mlp_forward(cfg, params, ep, ed, prec="f64")
    tp = torch.tensor(params, requires_grad=True)
    trd, trr = torch_spec.mlp(cfg, shapes, tp, torch.tensor(ep), torch.tensor(ed))
    np.testing.assert_allclose(rd, trd.detach().numpy(), atol=1e-12)
    np.testing.assert_allclose(rr, trr.detach().numpy(), atol=1e-12)
    gd, gr = rng.normal(size=M), rng.normal(size=(M, 3))
    ((trd * torch.tensor(gd)).sum() + (trr * torch.tensor(gr)).sum()).backward()
    g = orc.mlp_backward(cfg, params, ep, ed, acts, gd, gr, prec="f64")
    np.testing.assert_allclose(g, tp.grad.numpy(), atol=1e-11)
    rd32, rr32, acts32 = orc.mlp_forward(cfg, params, ep, ed, prec="f32")
    np.testing.assert_allclose(rd32, rd, rtol=1e-4, atol=1e-5)
    g32 = orc.mlp_backward(cfg, params, ep, ed, acts32, gd, gr, prec="f32")
    assert np.abs(g32 - g).max() <= 1e-4 * np.abs(g).max()


@pytest.mark.parametrize("bias,pad", [(0.0, 0.0), (-1.0, 0.001)])
def test_train_gradient_vs_autograd(bias, pad):
    """Whole step (SN/MipNerfModel.cs:99-200): fp64 oracle == torch-fp64 autograd of Appendix B."""
    cfg = small_cfg(density_bias=bias, rgb_padding=pad)
    shapes = orc.layer_shapes(cfg)
    R, S = 12, cfg.n_samples
    rays, pix = orc.synthetic_rays(R, width=100, height=100)
    rays["loss_mults"] = np.random.default_rng(6).uniform(0.5, 1.5, R).astype(np.float32)
    params = orc.init_params(cfg, 7)
    u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, R, S + 1) for lv in range(2)])
    o64 = orc.train_gradient(cfg, params, rays, pix, u, prec="f64")
    tp = torch.tensor(params.astype(np.float64), requires_grad=True)
    tr = {k: torch.tensor(v.astype(np.float64)) for k, v in rays.items()}
    loss, comps = torch_spec.total_loss(cfg, shapes, tp, tr, torch.tensor(pix.astype(np.float64)),
                                        [torch.tensor(o64["t_vals"][lv]) for lv in range(2)])
    loss.backward()
    assert np.isclose(o64["total_loss"], loss.item(), rtol=1e-10)
    np.testing.assert_allclose(o64["comp_rgb"][1], comps[1].detach().numpy(), atol=1e-11)
    scale = np.abs(tp.grad.numpy()).max()
    assert np.abs(o64["grads"] - tp.grad.numpy()).max() <= 1e-9 * scale
    # fp32 build tracks its fp64 shadow within the fp32-path tolerance of BASELINE.md (1e-4 of scale)
    o32 = orc.train_gradient(cfg, params, rays, pix, u, prec="f32")
    assert np.abs(o32["grads"] - o64["grads"]).max() <= 1e-4 * scale
    np.testing.assert_allclose(o32["comp_rgb"], o64["comp_rgb"], rtol=1e-4, atol=1e-5)
    # level-1 samples are sorted and inside [near, far]
    t1 = o64["t_vals"][1]
    assert np.all(np.diff(t1, axis=1) >= 0) and t1.min() >= 2.0 - 1e-6 and t1.max() <= 6.0 + 1e-6


def test_train_gradient_deterministic_across_threads():
    cfg = small_cfg()
    R, S = 16, cfg.n_samples
    rays, pix = orc.synthetic_rays(R, width=100, height=100)
    params = orc.init_params(cfg, 7)
    u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, R, S + 1) for lv in range(2)])
    n = orc.max_threads()
    try:
        orc.set_threads(1)
        a = orc.train_gradient(cfg, params, rays, pix, u, prec="f32")
        orc.set_threads(max(2, n))
        b = orc.train_gradient(cfg, params, rays, pix, u, prec="f32")
    finally:
        orc.set_threads(n)
    np.testing.assert_array_equal(a["comp_rgb"], b["comp_rgb"])  # per-ray work is order-free
    assert np.abs(a["grads"] - b["grads"]).max() <= 1e-5 * np.abs(a["grads"]).max()


def test_ssim_oracle_against_scipy_and_identities():
    """SN/MipHelpers.cs:688-757.  Independent check of the restatement: the same definition written with scipy's
    correlate2d (zero-padded `same` window of the normalised 11x11 Gaussian) in fp64; identities: SSIM(a, a) = 1,
    symmetric in its arguments, < 1 and decreasing with added noise."""
    from scipy.signal import correlate2d

    rng = np.random.default_rng(0)
    H, W = 37, 53  # not multiples of anything: borders and odd sizes
    a = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.05, a.shape), 0, 1).astype(np.float32)
    mean, m = orc.ssim(a, b)
    x = np.arange(11) - 5
    filt = np.exp(-(x[:, None] ** 2 + x[None, :] ** 2) / (2 * 1.5 ** 2))
    filt /= filt.sum()

    def conv(img):
        return np.stack([correlate2d(img[..., c].astype(np.float64), filt, mode="same", boundary="fill") for c in range(3)], -1)

    mu0, mu1 = conv(a), conv(b)
    s00 = np.maximum(conv(a.astype(np.float64) ** 2) - mu0 ** 2, 0)
    s11 = np.maximum(conv(b.astype(np.float64) ** 2) - mu1 ** 2, 0)
    s01 = np.maximum(conv(a.astype(np.float64) * b) - mu0 * mu1, 0)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ref = ((2 * mu0 * mu1 + c1) * (2 * s01 + c2)) / ((mu0 ** 2 + mu1 ** 2 + c1) * (s00 + s11 + c2))
    np.testing.assert_allclose(m, ref, atol=2e-4)  # fp32 window sums vs fp64
    assert abs(mean - ref.mean()) <= 2e-5
    assert abs(orc.ssim(a, a)[0] - 1.0) <= 1e-5
    assert abs(orc.ssim(b, a)[0] - mean) <= 1e-7
    worse = np.clip(a + rng.normal(0, 0.2, a.shape), 0, 1).astype(np.float32)
    assert orc.ssim(a, worse)[0] < mean < 1.0


def _pose(seed=0):
    """camera on a radius-4 sphere looking at the origin: 3x4 row-major [R | t] (columns of R: right, up, back)"""
    rng = np.random.default_rng(seed)
    th, ph = rng.uniform(0, 2 * np.pi), rng.uniform(0.2 * np.pi, 0.5 * np.pi)
    cam = 4.0 * np.array([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)])
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    return np.concatenate([np.stack([right, up, -fwd], -1), cam[:, None]], 1).astype(np.float32)


def test_generate_rays_oracle_against_fp64_statement():
    """SN/Dataset.cs:111-176 restated (oracle.c) against the same formulas written with numpy in fp64: directions, origins,
    radii = |d(x) - d(x+1)| * 2 / sqrt(12); the reference's zero radius at the last column (edge_mode 0) and the
    left-neighbour variant (edge_mode 1); a pixel sub-range equals the slice of the whole view."""
    W, H = 37, 23
    c2w = _pose(1)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    R, t = c2w[:, :3].astype(np.float64), c2w[:, 3].astype(np.float64)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cam = np.stack([(xs - W * 0.5 + 0.5) / focal, -(ys - H * 0.5 + 0.5) / focal, -np.ones_like(xs, dtype=np.float64)], -1)
    d = cam @ R.T
    for mode in (0, 1):
        o = orc.generate_rays(c2w, focal, W, H, edge_mode=mode)
        np.testing.assert_allclose(o["directions"].reshape(H, W, 3), d, rtol=2e-6, atol=2e-7)
        np.testing.assert_array_equal(o["origins"], np.broadcast_to(c2w[:, 3], (W * H, 3)))
        rad = np.linalg.norm(d[:, :-1] - d[:, 1:], axis=-1) * 2 / np.sqrt(12.0)
        got = o["radii"].reshape(H, W)
        np.testing.assert_allclose(got[:, :-1], rad, rtol=2e-3)  # a difference of nearly equal fp32 directions
        if mode == 0:
            assert (got[:, -1] == 0).all()                        # SN/Dataset.cs:151: the last column differences with itself
        else:
            np.testing.assert_allclose(got[:, -1], rad[:, -1], rtol=2e-3)
        assert (o["nears"] == 2).all() and (o["fars"] == 6).all()
        part = orc.generate_rays(c2w, focal, W, H, edge_mode=mode, first=100, n=333)
        for k in o:
            np.testing.assert_array_equal(part[k], o[k][100:433])


def test_resample_against_the_published_mipnerf_formulation():
    """Pin P6 for hierarchical sampling (the reference's CUDA version is defective, A-D9, and its C# cannot run here): the
    inversion of the piecewise-constant CDF written the way the mip-NeRF paper's public code states it — fully vectorised, with
    comparison masks and max / min reductions instead of the C#'s per-sample interval search (SN/MipHelpers.cs:774-851) —
    must give the same t-values as the oracle, for jittered and deterministic u, for peaked, flat and all-zero weights."""
    R, S = 6, 32
    rng = np.random.default_rng(7)
    t = np.sort(rng.uniform(2, 6, (R, S + 1)), 1)
    w = rng.uniform(0, 1, (R, S)) ** 6
    w[1] = 0.0
    w[2] = 1.0
    w[3, :] = 0.0
    w[3, 17] = 1.0
    u01 = orc.sampling_uniforms(5, 3, 1, 0, R, S + 1).astype(np.float64)
    pad_ = 0.01

    def published(randomized):
        wp = np.concatenate([w[:, :1], w, w[:, -1:]], 1)                       # blur-pool (mip-NeRF resample_along_rays)
        wmax = np.maximum(wp[:, :-1], wp[:, 1:])
        wb = 0.5 * (wmax[:, :-1] + wmax[:, 1:]) + pad_
        wsum = wb.sum(-1, keepdims=True)
        padding = np.maximum(0, 1e-5 - wsum)
        wb = wb + padding / S
        wsum = wsum + padding
        pdf = wb / wsum
        cdf = np.minimum(1, np.cumsum(pdf[:, :-1], -1))
        cdf = np.concatenate([np.zeros((R, 1)), cdf, np.ones((R, 1))], -1)     # [R, S+1]
        ns = S + 1
        if randomized:
            s = 1.0 / ns
            u = np.arange(ns) * s + u01 * (s - 1e-7)
            u = np.minimum(u, 1.0 - 1e-7)
        else:
            u = np.broadcast_to(np.linspace(0.0, 1.0 - 1.1920929e-7, ns), (R, ns))
        mask = u[:, None, :] >= cdf[:, :, None]                                  # [R, bins+1, samples]

        def find_interval(x):
            x0 = np.max(np.where(mask, x[:, :, None], x[:, :1, None]), -2)
            x1 = np.min(np.where(~mask, x[:, :, None], x[:, -1:, None]), -2)
            return x0, x1

        b0, b1 = find_interval(t)
        c0, c1 = find_interval(cdf)
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.clip(np.nan_to_num((u - c0) / (c1 - c0), nan=0.0, posinf=0.0, neginf=0.0), 0, 1)
        return b0 + frac * (b1 - b0)

    for randomized in (1, 0):
        got = orc.resample_t_vals(t, w, u01.astype(np.float32), pad_, randomized, prec="f64")
        np.testing.assert_allclose(got, published(randomized), rtol=0, atol=1e-9)
        assert np.all(np.diff(got, axis=1) >= -1e-12)
