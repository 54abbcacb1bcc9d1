"""-m gpu: per-stage parity of the CUDA kernels (through the C ABI) against the CPU oracle, same seeded inputs.

Tolerances (stated per BASELINE.md): index/ordering work (sampling) bit-exact; fp32 stages
|x - x_oracle| <= 1e-6 + 1e-4*|x_oracle| elementwise unless a test says why it is looser."""
import ctypes as C

import numpy as np
import pytest
import torch

import nerf_or_nothing_b200 as nb
from oracle import oracle as orc
from tests.gpu_util import call, dev, empty, host, ptr, rel_err, zeros

pytestmark = pytest.mark.gpu


def _rays(R, seed=0):
    rays, _ = orc.synthetic_rays(R, width=100, height=100, seed=2024 + seed)
    return rays


@pytest.mark.parametrize("R,S", [(1, 32), (7, 64), (130, 128), (33, 256)])
def test_sample_t_vals_bit_exact(R, S):
    rays = _rays(R)
    u = orc.sampling_uniforms(5, 1, 0, 0, R, S + 1)
    t = empty(R, S + 1)
    call("nerf_get_sample_t_vals", ptr(dev(rays["nears"])), ptr(dev(rays["fars"])), ptr(dev(u)), R, S, 1, ptr(t))
    np.testing.assert_array_equal(host(t), orc.sample_t_vals(rays["nears"], rays["fars"], u, S))
    call("nerf_get_sample_t_vals", ptr(dev(rays["nears"])), ptr(dev(rays["fars"])), None, R, S, 0, ptr(t))
    np.testing.assert_array_equal(host(t), orc.sample_t_vals(rays["nears"], rays["fars"], None, S, randomized=0))


@pytest.mark.parametrize("R,S", [(1, 32), (9, 64), (257, 128), (17, 256)])
def test_resample_t_vals_bit_exact(R, S):
    rng = np.random.default_rng(R * S)
    t = np.sort(rng.uniform(2, 6, (R, S + 1)).astype(np.float32), 1)
    w = (rng.uniform(0, 1, (R, S)) ** 8).astype(np.float32)
    w[0] = 0  # all-zero histogram: padding-only branch
    if R > 2:
        w[1] = 0
        w[1, S // 2] = 1.0  # a single spike: zero-width cdf steps
    u = orc.sampling_uniforms(7, 0, 1, 0, R, S + 1)
    out = empty(R, S + 1)
    for randomized in (1, 0):
        call("nerf_get_resampled_t_vals", ptr(dev(t)), ptr(dev(w)), ptr(dev(u)), R, S, 0.01, randomized, ptr(out))
        ref = orc.resample_t_vals(t, w, u, 0.01, randomized)
        got = host(out)
        np.testing.assert_array_equal(got, ref)
        assert np.all(np.diff(got, axis=1) >= 0)


@pytest.mark.parametrize("R,S", [(5, 64), (300, 128)])
def test_cast_rays_and_encode(R, S):
    rays = _rays(R, 1)
    u = orc.sampling_uniforms(3, 0, 0, 0, R, S + 1)
    t = orc.sample_t_vals(rays["nears"], rays["fars"], u, S)
    mean_o, cov_o = orc.cast_rays(t, rays["origins"], rays["directions"], rays["radii"])
    means, covs = empty(R, S, 3), empty(R, S, 3)
    call("nerf_cast_rays", ptr(dev(t)), ptr(dev(rays["origins"])), ptr(dev(rays["directions"])), ptr(means), ptr(covs),
         ptr(dev(rays["radii"])), R, S)
    # explicitly rounded ops in the oracle's order -> bit-identical Gaussians
    np.testing.assert_array_equal(host(means), mean_o)
    np.testing.assert_array_equal(host(covs), cov_o)
    enc_pos, enc_dir = empty(R * S, 96), empty(R * S, 27)
    call("nerf_encode_input_data", ptr(means), ptr(covs), ptr(dev(rays["directions"])), ptr(enc_pos), ptr(enc_dir), R, S, 16, 4)
    ep_o = orc.encode_position(mean_o, cov_o, 16)
    # same fp32 arguments; only libm (expf/sincosf) ulp differences remain
    np.testing.assert_allclose(host(enc_pos), ep_o, rtol=1e-5, atol=2e-7)
    ed_o = np.repeat(orc.encode_direction(rays["directions"], 4), S, axis=0)
    np.testing.assert_allclose(host(enc_dir), ed_o, rtol=1e-6, atol=2e-7)
    # fp64 shadow on the same fp32 inputs: fp32 IPE is exact up to ~1 ulp of the (exactly scaled) argument
    ep64 = orc.encode_position(mean_o.astype(np.float64), cov_o.astype(np.float64), 16, prec="f64")
    np.testing.assert_allclose(host(enc_pos), ep64, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("R,S", [(3, 32), (100, 64), (1000, 128), (64, 256)])
@pytest.mark.parametrize("white", [1, 0])
def test_volumetric_rendering_fwd_bwd(R, S, white):
    rng = np.random.default_rng(R + S)
    rays = _rays(R, 2)
    rgb = rng.uniform(0, 1, (R, S, 3)).astype(np.float32)
    den = (rng.uniform(0, 1, (R, S)) ** 4 * 40).astype(np.float32)
    den[0] = 0  # empty ray: acc = 0 -> depth clamps to t_S
    t = np.sort(rng.uniform(2, 6, (R, S + 1)).astype(np.float32), 1)
    d = rays["directions"]
    comp, depth, acc, w = empty(R, 3), empty(R), empty(R), empty(R, S)
    call("nerf_volumetric_rendering", ptr(dev(rgb)), ptr(dev(den)), ptr(dev(t)), ptr(dev(d)), ptr(comp), ptr(depth), ptr(acc),
         ptr(w), R, S, white)
    o32 = orc.volumetric_rendering(rgb, den, t, d, white)
    o64 = orc.volumetric_rendering(rgb, den, t, d, white, prec="f64")
    for got, key in ((comp, "comp_rgb"), (acc, "acc"), (w, "weights"), (depth, "depth")):
        np.testing.assert_allclose(host(got), o32[key], rtol=1e-4, atol=1e-6, err_msg=key)
        np.testing.assert_allclose(host(got), o64[key], rtol=1e-4, atol=1e-6, err_msg=key + " (f64)")
    assert host(depth)[0] == t[0, -1]
    # size-independent identity: sum(w) + T_end = 1
    T_end = o64["transmittance"][:, -1] * (1 - o64["alpha"][:, -1])
    np.testing.assert_allclose(host(acc) + T_end, 1.0, atol=2e-6)
    g = rng.normal(size=(R, 3)).astype(np.float32)
    for mode in (0, 1):
        d_rgb, d_den = empty(R, S, 3), empty(R, S)
        call("nerf_volumetric_rendering_gradient", ptr(dev(g)), ptr(dev(rgb)), ptr(dev(den)), ptr(dev(t)), ptr(dev(d)),
             ptr(d_rgb), ptr(d_den), R, S, white, mode)
        r_rgb, r_den = orc.volumetric_rendering_gradient(g, rgb, den, t, d, white, mode, prec="f64")
        np.testing.assert_allclose(host(d_rgb), r_rgb, rtol=1e-4, atol=1e-6)
        # d_density sums ~S signed terms: tolerance on the ray's gradient scale
        scale = np.abs(r_den).max(axis=1, keepdims=True) + 1e-6
        assert np.abs(host(d_den) - r_den).max() <= 1e-4 * scale.max()
        assert (np.abs(host(d_den) - r_den) / scale).max() <= 2e-4


@pytest.mark.parametrize("R,S", [(5, 32), (77, 64), (1001, 128), (40, 256)])
@pytest.mark.parametrize("bias,pad", [(0.0, 0.0), (-1.0, 0.001)])
def test_volumetric_rendering_fused_activations(R, S, bias, pad):
    """The form the model runs (SN/MipNerfModel.cs:81-83, 184-189 fused into the compositing kernels), through the
    asynchronous entry on a non-default stream: raw head outputs in, gradients w.r.t. the raw head outputs out."""
    rng = np.random.default_rng(7 * R + S)
    rays = _rays(R, 3)
    raw_rgb = rng.normal(0, 2.5, (R, S, 3)).astype(np.float32)
    raw_den = rng.normal(-1, 3, (R, S)).astype(np.float32)
    raw_den[0] = -60.0   # softplus underflows: empty ray
    raw_rgb[1] = 30.0    # saturated sigmoid
    t = np.sort(rng.uniform(2, 6, (R, S + 1)).astype(np.float32), 1)
    d = rays["directions"]
    g = rng.normal(size=(R, 3)).astype(np.float32)
    cfg = orc.default_config(density_bias=bias, rgb_padding=pad)
    den64, rgb64 = orc.output_activations(cfg, raw_den, raw_rgb, prec="f64")
    den64, rgb64 = den64.reshape(R, S), rgb64.reshape(R, S, 3)
    o64 = orc.volumetric_rendering(rgb64, den64, t, d, 1, prec="f64")
    comp, depth, acc, w = empty(R, 3), empty(R), empty(R), empty(R, S)
    st = torch.cuda.Stream()
    sp = C.c_void_p(st.cuda_stream)
    args = [dev(raw_rgb), dev(raw_den), dev(t), dev(d), dev(g)]
    torch.cuda.synchronize()
    call("nerf_volumetric_rendering_async", ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(args[3]), ptr(comp), ptr(depth),
         ptr(acc), ptr(w), R, S, 1, 1, bias, pad, sp)
    st.synchronize()
    for got, key in ((comp, "comp_rgb"), (acc, "acc"), (w, "weights")):
        np.testing.assert_allclose(host(got), o64[key], rtol=1e-4, atol=2e-6, err_msg=key)
    # ray 0: fp32 softplus(-60) is exactly 0 (as in the reference's fp32 arithmetic) -> acc = 0 -> depth clamps to t_S
    np.testing.assert_allclose(host(depth)[1:], o64["depth"][1:], rtol=1e-4, atol=2e-6, err_msg="depth")
    assert host(depth)[0] == t[0, -1] and host(acc)[0] == 0.0
    for mode in (0, 1):
        d_rgb, d_den = empty(R, S, 3), empty(R, S)
        call("nerf_volumetric_rendering_gradient_async", ptr(args[4]), ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(args[3]),
             ptr(d_rgb), ptr(d_den), R, S, 1, mode, 1, bias, pad, sp)
        st.synchronize()
        a_rgb, a_den = orc.volumetric_rendering_gradient(g, rgb64, den64, t, d, 1, mode, prec="f64")
        r_den, r_rgb = orc.output_activations_grad(cfg, raw_den, raw_rgb, a_den, a_rgb, prec="f64")
        r_den, r_rgb = r_den.reshape(R, S), r_rgb.reshape(R, S, 3)
        np.testing.assert_allclose(host(d_rgb), r_rgb, rtol=1e-4, atol=1e-6)
        # d_density sums ~S signed terms of size |g| each: tolerance on the ray's gradient scale, floored at 1e-4 |g| so that
        # the saturated white ray (true gradient ~0 by cancellation of g.c against sum(g)) is judged against fp32 rounding
        scale = np.abs(r_den).max(axis=1, keepdims=True) + 1e-4 * np.abs(g).sum(axis=1, keepdims=True) + 1e-6
        assert (np.abs(host(d_den) - r_den) / scale).max() <= 2e-4


def test_output_gradient_and_adam():
    rng = np.random.default_rng(11)
    R = 777
    comp, pix = rng.uniform(0, 1, (R, 3)).astype(np.float32), rng.uniform(0, 1, (R, 3)).astype(np.float32)
    lm = rng.uniform(0.5, 2, R).astype(np.float32)
    g = empty(R, 3)
    call("nerf_get_output_gradient", ptr(dev(comp)), ptr(dev(pix)), ptr(dev(lm)), ptr(g), float(lm.sum()), 0.1, R)
    np.testing.assert_allclose(host(g), orc.output_gradient(comp, pix, lm, float(lm.sum()), 0.1), rtol=1e-5, atol=1e-9)
    n = 546948 + 3  # flat parameter count of the 8x256 net (+3: exercises the non-multiple-of-4 tail)
    p, gr = rng.normal(size=n).astype(np.float32), (rng.normal(size=n) * 1e-3).astype(np.float32)
    m, v = (rng.normal(size=n) * 1e-3).astype(np.float32), (rng.uniform(0, 1e-5, n)).astype(np.float32)
    for eps_mode in (0, 1):
        dp, dm, dv = dev(p), dev(m), dev(v)
        it = 5
        inv1, inv2 = 1 / (1 - np.float32(0.9) ** it), 1 / (1 - np.float32(0.999) ** it)
        call("nerf_adam_optimizer_step", ptr(dp), ptr(dev(gr)), ptr(dm), ptr(dv), 1e-3, 0.9, 0.999, float(inv1), float(inv2), n, eps_mode)
        p64, m64, v64 = orc.adam_step(p, gr, m, v, 1e-3, it, eps_mode, prec="f64")
        np.testing.assert_allclose(host(dm), m64, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(host(dv), v64, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(host(dp) - p, p64 - p, rtol=1e-3, atol=5e-7)  # the update itself (p ~ 4: fp32 ulp(p) = 4.8e-7)
        np.testing.assert_allclose(host(dp), p64, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("act", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(256, 96, 0), (256, 256, 96), (128, 256, 27), (1, 256, 0), (3, 128, 0)])
def test_apply_and_backpropagate_layer(act, shape):
    n, ka, kb = shape
    if n <= 4 and act == 0:
        pytest.skip("heads use sigmoid/softplus/identity")
    rng = np.random.default_rng(n * 7 + ka + kb + act)
    M = 1000  # not a multiple of the 128-row tile
    xa = rng.normal(size=(M, ka)).astype(np.float32)
    xb = rng.normal(size=(M, kb)).astype(np.float32) if kb else None
    W = (rng.normal(size=(n, ka + kb)) / np.sqrt(ka + kb)).astype(np.float32)
    b = rng.normal(size=n).astype(np.float32) * 0.1
    x = np.concatenate([xa, xb], 1) if kb else xa
    z64 = x.astype(np.float64) @ W.astype(np.float64).T + b
    f = {0: lambda z: np.maximum(z, 0), 1: lambda z: 1 / (1 + np.exp(-z)), 2: lambda z: np.log1p(np.exp(z)), 3: lambda z: z}[act]
    df = {0: lambda z: (z > 0) * 1.0, 1: lambda z: f(z) * (1 - f(z)), 2: lambda z: 1 / (1 + np.exp(-z)), 3: lambda z: np.ones_like(z)}[act]
    out, z = empty(M, n), empty(M, n)
    dxa, dxb, dW, db = dev(xa), (dev(xb) if kb else None), dev(W), dev(b)
    call("nerf_apply_layer", ptr(dxa), ptr(dxb), ptr(dW), ptr(db), ptr(out), ptr(z), M, n, ka, kb, act)
    np.testing.assert_allclose(host(z), z64, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(host(out), f(z64), rtol=1e-4, atol=2e-5)
    dy = rng.normal(size=(M, n)).astype(np.float32)
    gin, gW, gb = zeros(M, ka), zeros(n, ka + kb), zeros(n)
    for _ in range(2):  # gradients accumulate like the reference's atomicAdd (.cu:105-110)
        call("nerf_backpropagate_layer", ptr(dxa), ptr(dxb), ptr(dW), ptr(z), ptr(dev(dy)), ptr(gin), ptr(gW), ptr(gb), M, n, ka, kb, act)
    dz = dy.astype(np.float64) * df(host(z).astype(np.float64))
    assert rel_err(host(gW), 2 * dz.T @ x.astype(np.float64)) <= 1e-4
    assert rel_err(host(gb), 2 * dz.sum(0)) <= 1e-4
    assert rel_err(host(gin), 2 * (dz @ W.astype(np.float64))[:, :ka]) <= 1e-4
