"""-m gpu: the tcgen05/TMEM MLP engine (NERF_PRECISION_FP32_TC = bf16x3 split, NERF_PRECISION_BF16_TC) against the
fp64 oracle.  Tolerances are BASELINE.json's: <= 1e-4 relative (to the tensor scale) for the fp32-accurate mode,
<= 2e-2 for bf16."""
import numpy as np
import pytest

import nerf_or_nothing_b200 as nb
from oracle import oracle as orc
from tests.gpu_util import batch, configs_pair, dev, from_ptr, rel_err

pytestmark = pytest.mark.gpu

TOL = {"fp32_tc": 1e-4, "bf16": 2e-2}
NET = dict(n_samples=64)  # 8x256 + 1x128 view branch, 64+64 samples (BASELINE config 1 shape)
SMALL = dict(n_samples=32, net_depth=4, net_width=64, net_depth_condition=2, net_width_condition=64, skip_layer=2,
             deg_point=8, deg_view=2)


def _model(R, precision, **kw):
    ncfg, ocfg = configs_pair(n_rays=R, precision=nb.PRECISIONS[precision], **kw)
    return nb.AcceleratedMipNeRF(ncfg), ncfg, ocfg


def _params_with_biases(ocfg, seed=5):
    rng = np.random.default_rng(seed)
    params = orc.init_params(ocfg, 7)
    nb_ = sum(orc.layer_shapes(ocfg)[0])
    params[-nb_:] = rng.normal(size=nb_).astype(np.float32) * 0.1
    return params


@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
@pytest.mark.parametrize("cfgkw,R", [(NET, 16), (NET, 3), (SMALL, 5)], ids=["8x256-M1024", "8x256-M192", "small-M160"])
def test_tc_mlp_forward(precision, cfgkw, R):
    m, ncfg, ocfg = _model(R, precision, **cfgkw)
    S = ncfg.n_samples
    M = R * S
    rng = np.random.default_rng(3)
    P, Dd = 6 * ncfg.deg_point, 3 + 6 * ncfg.deg_view
    params = _params_with_biases(ocfg)
    m.set_params(params)
    ep = rng.uniform(-1, 1, (M, P)).astype(np.float32)
    ed = rng.uniform(-1, 1, (M, Dd)).astype(np.float32)
    d_ptr, r_ptr = m.mlp.get_output(dev(ep), dev(ed), 0, R)
    rd64, rr64, _ = orc.mlp_forward(ocfg, params, ep, ed, prec="f64")
    den64, rgb64 = orc.output_activations(ocfg, rd64, rr64, prec="f64")
    e_d, e_r = rel_err(from_ptr(d_ptr, (M,)), den64), rel_err(from_ptr(r_ptr, (M, 3)), rgb64)
    print(f"{precision}: density rel err {e_d:.2e}, rgb rel err {e_r:.2e}")
    assert e_d <= TOL[precision] and e_r <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
@pytest.mark.parametrize("cfgkw,R", [(NET, 16), (NET, 3), (SMALL, 5)], ids=["8x256-M1024", "8x256-M192", "small-M160"])
def test_tc_mlp_backward(precision, cfgkw, R):
    m, ncfg, ocfg = _model(R, precision, **cfgkw)
    S = ncfg.n_samples
    M = R * S
    rng = np.random.default_rng(4)
    P, Dd = 6 * ncfg.deg_point, 3 + 6 * ncfg.deg_view
    params = _params_with_biases(ocfg)
    m.set_params(params)
    ep = rng.uniform(-1, 1, (M, P)).astype(np.float32)
    ed = rng.uniform(-1, 1, (M, Dd)).astype(np.float32)
    m.mlp.get_output(dev(ep), dev(ed), 1, R)
    cg, dg = rng.normal(size=(M, 3)).astype(np.float32), rng.normal(size=M).astype(np.float32)
    m.mlp.reset_gradients(1)
    m.mlp.get_gradient(dev(cg), dev(dg), 1)
    rd64, rr64, acts64 = orc.mlp_forward(ocfg, params, ep, ed, prec="f64")
    d_rd, d_rr = orc.output_activations_grad(ocfg, rd64, rr64, dg, cg, prec="f64")
    g64 = orc.mlp_backward(ocfg, params, ep, ed, acts64, d_rd, d_rr, prec="f64")
    got = m.get_gradients()
    sizes, off, worst = m.GetLayerSizes(), 0, (0.0, -1)
    for i, n in enumerate(sizes):
        e = rel_err(got[off:off + n], g64[off:off + n])
        worst = max(worst, (e, i))
        off += n
    print(f"{precision}: grad rel err {rel_err(got, g64):.2e}; worst tensor {worst[1]} at {worst[0]:.2e}")
    assert rel_err(got, g64) <= TOL[precision]
    assert worst[0] <= 3 * TOL[precision], f"tensor {worst[1]}"


@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
def test_tc_whole_step_and_training(precision):
    R = 64
    m, ncfg, ocfg = _model(R, precision, **NET)
    S = ncfg.n_samples
    rays, pix, u = batch(R, S)
    m.set_pixels(pix)
    m.set_sampling_uniforms(u)
    m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
    params = orc.init_params(ocfg, 7)
    o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64")
    tol = TOL[precision]
    for lv in range(2):
        out = m.level_outputs(lv)
        assert rel_err(from_ptr(out["comp_rgb"], (R, 3)), o64["comp_rgb"][lv]) <= tol
    per, total = m.get_loss()
    assert abs(total - o64["total_loss"]) <= tol * o64["total_loss"]
    e = rel_err(m.get_gradients(), o64["grads"])
    print(f"{precision}: whole-step grad rel err {e:.2e}")
    assert e <= tol
    # loss curve over 30 Adam steps within 1 % of the fp32 CUDA-core path on identical batches
    ref, _, _ = _model(R, "fp32", **NET)
    opt_a, opt_b = nb.AcceleratedAdamOptimizer(m.GetLayerSizes()), nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
    m.set_params(params)
    for step in range(1, 31):
        rays, pix, u = batch(R, S, seed=step, step=step)
        losses = []
        for mod, opt in ((m, opt_a), (ref, opt_b)):
            mod.set_sampling_uniforms(u)
            losses.append(mod.train_step(opt, rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                                         rays["loss_mults"], pix, 1e-3))
        assert abs(losses[0] - losses[1]) <= 1e-2 * losses[1], (step, losses)
