"""-m gpu: whole-path parity through the reference-shaped interface (AcceleratedMipNeRF / AcceleratedMLP /
AcceleratedAdamOptimizer / AcceleratedGradientCalculator / OutputRetriever) against the CPU oracle.

fp32 path tolerance: <= 1e-4 relative to the tensor scale against the fp64 shadow, on identical
inputs, weights (same Philox stream) and sampling uniforms."""
import numpy as np
import pytest
import torch

import nerf_or_nothing_b200 as nb
from oracle import oracle as orc
from tests.gpu_util import batch, configs_pair, dev, empty, from_ptr, host, ptr, rel_err

pytestmark = pytest.mark.gpu

SMALL = dict(n_samples=32, net_depth=4, net_width=64, net_depth_condition=1, net_width_condition=32, skip_layer=2,
             deg_point=8, deg_view=2)
REF = dict(n_samples=64)  # BASELINE config 1: 8x256 net, 64+64 samples


def _model(R, precision="fp32", **kw):
    ncfg, ocfg = configs_pair(n_rays=R, precision=nb.PRECISIONS[precision], **kw)
    return nb.AcceleratedMipNeRF(ncfg), ncfg, ocfg


def test_init_matches_oracle_stream_and_layer_sizes():
    m, ncfg, ocfg = _model(64)
    assert m.GetLayerSizes() == orc.layer_sizes(ocfg)  # ANU/AcceleratedMLP.cpp:131-154
    assert m.num_params == 546948
    p = m.get_params()
    np.testing.assert_allclose(p, orc.init_params(ocfg, 7), rtol=1e-6, atol=1e-8)
    ptrs = m.mlp.allParams
    sizes = m.GetLayerSizes()
    assert all(ptrs[i + 1] - ptrs[i] == 4 * sizes[i] for i in range(len(sizes) - 1))  # views of one flat buffer


@pytest.mark.parametrize("cfgkw", [SMALL, REF], ids=["small", "8x256"])
def test_mlp_get_output_and_gradient(cfgkw):
    R = 24
    m, ncfg, ocfg = _model(R, **cfgkw)
    S = ncfg.n_samples
    M = R * S
    rng = np.random.default_rng(3)
    P, Dd = 6 * ncfg.deg_point, 3 + 6 * ncfg.deg_view
    params = orc.init_params(ocfg, 7)
    params[-sum(orc.layer_shapes(ocfg)[0]):] = rng.normal(size=sum(orc.layer_shapes(ocfg)[0])).astype(np.float32) * 0.1
    m.set_params(params)
    ep = rng.uniform(-1, 1, (M, P)).astype(np.float32)
    ed = rng.uniform(-1, 1, (M, Dd)).astype(np.float32)
    d_ptr, r_ptr = m.mlp.get_output(dev(ep), dev(ed), 1, R)
    rd64, rr64, acts64 = orc.mlp_forward(ocfg, params, ep, ed, prec="f64")
    den64, rgb64 = orc.output_activations(ocfg, rd64, rr64, prec="f64")
    assert rel_err(from_ptr(d_ptr, (M,)), den64) <= 1e-4
    assert rel_err(from_ptr(r_ptr, (M, 3)), rgb64) <= 1e-4
    cg, dg = rng.normal(size=(M, 3)).astype(np.float32), rng.normal(size=M).astype(np.float32)
    m.mlp.reset_gradients(1)
    m.mlp.get_gradient(dev(cg), dev(dg), 1)
    d_rd, d_rr = orc.output_activations_grad(ocfg, rd64, rr64, dg, cg, prec="f64")
    g64 = orc.mlp_backward(ocfg, params, ep, ed, acts64, d_rd, d_rr, prec="f64")
    assert rel_err(m.get_gradients(), g64) <= 1e-4
    # per-tensor check too: every one of the 22 gradients within 1e-4 of its own scale (biases included)
    sizes, off = m.GetLayerSizes(), 0
    got = m.get_gradients()
    for i, n in enumerate(sizes):
        assert rel_err(got[off:off + n], g64[off:off + n]) <= 2e-4, f"tensor {i}"
        off += n


@pytest.mark.parametrize("cfgkw,R", [(SMALL, 100), (REF, 64)], ids=["small", "8x256"])
@pytest.mark.parametrize("bias,pad", [(0.0, 0.0), (-1.0, 0.001)])
def test_get_gradient_whole_step(cfgkw, R, bias, pad):
    """GetGradient (ANU/AcceleratedMipNeRF.cpp:52-144) == oracle whole step (SN/MipNerfModel.cs:99-200)."""
    m, ncfg, ocfg = _model(R, density_bias=bias, rgb_padding=pad, **cfgkw)
    S = ncfg.n_samples
    rays, pix, u = batch(R, S)
    rays["loss_mults"] = np.random.default_rng(1).uniform(0.5, 1.5, R).astype(np.float32)
    m.set_pixels(pix)
    m.set_sampling_uniforms(u)
    m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
    params = orc.init_params(ocfg, 7)
    o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64")
    o32 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f32")
    lo = m.level_outputs(0)
    # level 0: identical t (bit-exact sampling from the same uniforms)
    np.testing.assert_array_equal(from_ptr(lo["t_vals"], (R, S + 1)), o32["t_vals"][0])
    for lv in range(2):
        out = m.level_outputs(lv)
        np.testing.assert_allclose(from_ptr(out["comp_rgb"], (R, 3)), o64["comp_rgb"][lv], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(from_ptr(out["acc"], (R,)), o64["acc"][lv], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(from_ptr(out["weights"], (R, S)), o64["weights"][lv], rtol=1e-3, atol=1e-5)
    # level 1 t-values come from level-0 weights: continuous in them, so close but not bit-equal
    np.testing.assert_allclose(from_ptr(m.level_outputs(1)["t_vals"], (R, S + 1)), o64["t_vals"][1], rtol=1e-4, atol=1e-4)
    per, total = m.get_loss()
    np.testing.assert_allclose(per, o64["loss"], rtol=1e-4)
    assert abs(total - o64["total_loss"]) <= 1e-4 * o64["total_loss"]
    g = m.get_gradients()
    e_gpu, e_cpu32 = rel_err(g, o64["grads"]), rel_err(o32["grads"], o64["grads"])
    print(f"grad rel err vs fp64 oracle: gpu {e_gpu:.2e}, fp32 oracle {e_cpu32:.2e}")
    assert e_gpu <= 1e-4


def test_callback_path_with_gradient_calculator():
    """The reference's TrainStep wiring (SN/Program.cs:48-62): GetGradient with a host callback that calls
    AcceleratedGradientCalculator.get_output_gradient, then optimizer.step(allParams, grads, lr), then
    OutputRetriever.RetrieveOutput — gives the same step as the fused built-in path."""
    R = 64
    rays, pix, u = batch(R, 32)
    results = []
    for use_cb in (True, False):
        m, ncfg, ocfg = _model(R, **SMALL)
        opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
        calc = nb.AcceleratedGradientCalculator(R)
        m.set_sampling_uniforms(u)
        seen = {}
        if use_cb:
            def cb(inputptr, level, loss_mult_sum, loss_mults):
                seen[level] = inputptr
                return calc.get_output_gradient(inputptr, pix, loss_mults, loss_mult_sum, level)
            grad = m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                                 rays["loss_mults"], cb)
            assert sorted(seen) == [0, 1]
            out = nb.OutputRetriever.RetrieveOutput(seen[1], R)
            np.testing.assert_array_equal(out, from_ptr(m.level_outputs(1)["comp_rgb"], (R, 3)))
        else:
            m.set_pixels(pix)
            grad = m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                                 rays["loss_mults"])
        g = m.get_gradients()
        opt.step(m.mlp.allParams, grad, 1e-3)
        torch.cuda.synchronize()
        results.append((g, m.get_params()))
    np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(results[0][1], results[1][1], rtol=1e-6, atol=1e-9)
    p0 = orc.init_params(configs_pair(**SMALL)[1], 7)
    assert not np.allclose(results[0][1], p0)


def test_train_loop_tracks_oracle_loss_curve():
    """20 Adam steps on identical batches/uniforms: per-step loss within 1% of the fp32 oracle, final params close."""
    R, steps = 64, 20
    m, ncfg, ocfg = _model(R, **SMALL)
    S = ncfg.n_samples
    opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
    params = orc.init_params(ocfg, 7)
    mo, vo = np.zeros_like(params), np.zeros_like(params)
    for step in range(1, steps + 1):
        rays, pix, u = batch(R, S, seed=step, step=step)
        lr = 1e-3
        m.set_sampling_uniforms(u)
        loss = m.train_step(opt, rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                            rays["loss_mults"], pix, lr)
        o = orc.train_gradient(ocfg, params, rays, pix, u, prec="f32")
        params, mo, vo = orc.adam_step(params, o["grads"], mo, vo, lr, step, 0, prec="f32")
        assert abs(loss - o["total_loss"]) <= 1e-2 * o["total_loss"], (step, loss, o["total_loss"])
    assert rel_err(m.get_params(), params) <= 2e-3


def test_chunked_equals_unchunked_and_deterministic():
    R = 96
    rays, pix, u = batch(R, 32)
    grads = []
    for chunk in (0, 32, 0):
        m, ncfg, _ = _model(R, chunk_rays=chunk, **SMALL)
        m.set_pixels(pix)
        m.set_sampling_uniforms(u)
        m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
        grads.append((m.get_gradients(), from_ptr(m.level_outputs(1)["comp_rgb"], (R, 3)), m.get_loss()[1]))
    np.testing.assert_array_equal(grads[0][0], grads[2][0])  # bitwise reproducible (no atomics anywhere)
    np.testing.assert_array_equal(grads[0][1], grads[1][1])  # per-ray outputs do not depend on chunking
    assert rel_err(grads[1][0], grads[0][0]) <= 1e-5          # gradient sums only differ by association
    assert abs(grads[1][2] - grads[0][2]) <= 1e-5 * grads[0][2]


def test_render_matches_oracle_forward():
    """Evaluation is deterministic (SN/MipNerfModel.cs:36-97 renders with randomized = false): a handle configured for
    jittered TRAINING still renders the midpoint samples, twice the same bits, without touching its training RNG state."""
    R = 200  # > chunk: exercises the chunk loop
    for randomized in (1, 0):
        m, ncfg, ocfg = _model(64, randomized=randomized, **SMALL)
        S = ncfg.n_samples
        rays, pix, _ = batch(R, S)
        u = np.zeros((2, R, S + 1), np.float32)  # ignored: the oracle evaluates with randomized = 0
        m.set_sampling_uniforms(np.full((2, 64, S + 1), 0.75, np.float32))  # explicit training uniforms must not leak into a render
        rgb, depth, acc = m.render(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
        again = m.render(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
        for x, y in zip((rgb, depth, acc), again):
            np.testing.assert_array_equal(x, y)
        ocfg.randomized = 0
        o = orc.train_gradient(ocfg, orc.init_params(ocfg, 7), rays, pix, u, with_backward=False, prec="f64")
        np.testing.assert_allclose(rgb, o["comp_rgb"][1], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(acc, o["acc"][1], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(depth, o["depth"][1], rtol=1e-4, atol=1e-4)


def test_full_size_properties_config2():
    """BASELINE config 2 shape (4096 rays, 128+128 samples, 8x256): size-independent properties only."""
    R, S = 4096, 128
    m, ncfg, _ = _model(R)
    rays, pix, _ = batch(R, S, width=800)
    m.set_pixels(pix)
    m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
    g1 = m.get_gradients()
    assert np.isfinite(g1).all() and np.abs(g1).max() > 0
    for lv in range(2):
        out = m.level_outputs(lv)
        w = from_ptr(out["weights"], (R, S))
        acc = from_ptr(out["acc"], (R,))
        comp = from_ptr(out["comp_rgb"], (R, 3))
        t = from_ptr(out["t_vals"], (R, S + 1))
        assert np.all(w >= 0) and np.all(acc <= 1 + 1e-5)
        np.testing.assert_allclose(w.sum(1), acc, rtol=1e-4, atol=1e-6)
        assert np.all(comp >= -1e-5) and np.all(comp <= 1 + 1e-5)  # convex combination with a white background
        assert np.all(np.diff(t, axis=1) >= 0) and t.min() >= 2 - 1e-5 and t.max() <= 6 + 1e-5
    # linearity of the gradient in the loss multipliers: doubling lm leaves g unchanged (normalised by sum(lm))
    m.set_step(0)
    lm2 = rays["loss_mults"] * 2
    m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], lm2)
    assert rel_err(m.get_gradients(), g1) <= 1e-5
