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
# the narrow end of BASELINE.json configs[4] (4x128, condition width 64): runs through the same fused kernels as 8x256
NARROW = dict(n_samples=64, net_depth=4, net_width=128, net_depth_condition=1, net_width_condition=64, skip_layer=2)
NETS = {"8x256": NET, "4x128": NARROW}


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
@pytest.mark.parametrize("cfgkw,R", [(NET, 16), (NET, 3), (SMALL, 5), (NARROW, 16)], ids=["8x256-M1024", "8x256-M192", "small-M160", "4x128-M1024"])
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


def _stable_inputs(ocfg, params, M, P, Dd, margin, seed):
    """Draw samples whose every ReLU pre-activation is at least `margin` away from 0 (fp64): the gradient is
    discontinuous at z = 0, so parity of the backward ARITHMETIC is only defined where no unit changes side."""
    import torch

    from tests import torch_spec

    rng = np.random.default_rng(seed)
    shapes = orc.layer_shapes(ocfg)
    keep_p, keep_d, have = [], [], 0
    tp = torch.tensor(params.astype(np.float64))
    while have < M:
        ep = rng.uniform(-1, 1, (4 * M, P)).astype(np.float32)
        ed = rng.uniform(-1, 1, (4 * M, Dd)).astype(np.float32)
        mg = torch_spec.relu_margin(ocfg, shapes, tp, torch.tensor(ep.astype(np.float64)), torch.tensor(ed.astype(np.float64))).numpy()
        ok = mg > margin
        keep_p.append(ep[ok]); keep_d.append(ed[ok]); have += int(ok.sum())
    return np.concatenate(keep_p)[:M], np.concatenate(keep_d)[:M]


@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
@pytest.mark.parametrize("cfgkw,R", [(NET, 64), (NET, 3), (SMALL, 5), (NET, 300), (NARROW, 64)],
                         ids=["8x256-M4096", "8x256-M192", "small-M160", "8x256-M19200", "4x128-M4096"])  # M19200: 75 head partials -> wide reduction
def test_tc_mlp_backward(precision, cfgkw, R):
    """dgrad / wgrad GEMMs (MN-major operands, split reduction) against the fp64 oracle.
    fp32_tc: <= 1e-4 of each tensor's scale on ReLU-stable samples (margin 2e-4 >> the 3e-6 forward error).
    bf16: the forward error (~2e-3) itself moves ~0.2 % of the ReLU masks, which bounds gradient agreement at
    ~sqrt(2*eps) ~ 6e-2 whatever the backward arithmetic; checked as relative L2 <= 0.15 per tensor."""
    m, ncfg, ocfg = _model(R, precision, **cfgkw)
    S = ncfg.n_samples
    M = R * S
    rng = np.random.default_rng(4)
    P, Dd = 6 * ncfg.deg_point, 3 + 6 * ncfg.deg_view
    params = _params_with_biases(ocfg)
    m.set_params(params)
    ep, ed = _stable_inputs(ocfg, params, M, P, Dd, 2e-4 if precision == "fp32_tc" else 0.0, 4)
    m.mlp.get_output(dev(ep), dev(ed), 1, R)
    cg, dg = rng.normal(size=(M, 3)).astype(np.float32), rng.normal(size=M).astype(np.float32)
    m.mlp.reset_gradients(1)
    m.mlp.get_gradient(dev(cg), dev(dg), 1)
    rd64, rr64, acts64 = orc.mlp_forward(ocfg, params, ep, ed, prec="f64")
    d_rd, d_rr = orc.output_activations_grad(ocfg, rd64, rr64, dg, cg, prec="f64")
    g64 = orc.mlp_backward(ocfg, params, ep, ed, acts64, d_rd, d_rr, prec="f64")
    got = m.get_gradients()
    sizes, off, worst, worst_l2 = m.GetLayerSizes(), 0, (0.0, -1), (0.0, -1)
    for i, n in enumerate(sizes):
        a, b = got[off:off + n].astype(np.float64), g64[off:off + n]
        worst = max(worst, (rel_err(a, b), i))
        worst_l2 = max(worst_l2, (float(np.linalg.norm(a - b) / np.linalg.norm(b)), i))
        off += n
    print(f"{precision} M={M}: worst tensor max-norm err {worst[0]:.2e} (#{worst[1]}), worst rel-L2 {worst_l2[0]:.2e} (#{worst_l2[1]})")
    if precision == "fp32_tc":
        assert worst[0] <= 1e-4, f"tensor {worst[1]}"
    else:
        assert worst_l2[0] <= 0.15, f"tensor {worst_l2[1]}"


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
    # whole-step gradients include the ReLU-kink effect (a forward error eps moves a fraction ~eps of the masks, worth
    # ~sqrt(2 eps) of gradient agreement): 1e-3 for the 3e-6 forward error of bf16x3, 0.15 (relative L2) for bf16
    g, g64 = m.get_gradients().astype(np.float64), o64["grads"]
    e, e2 = rel_err(g, g64), float(np.linalg.norm(g - g64) / np.linalg.norm(g64))
    print(f"{precision}: whole-step grad max-norm err {e:.2e}, rel-L2 {e2:.2e}")
    assert (e <= 1e-3) if precision == "fp32_tc" else (e2 <= 0.15)
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


LAYERED = nb.FLAG_NO_FUSED_FORWARD | nb.FLAG_NO_FUSED_TRAIN_FORWARD | nb.FLAG_NO_FUSED_DGRAD
# engine variants that must agree with the layer-by-layer kernels (nerf_config.engine_flags): the shipped one (cast_rays + IPE
# + direction PE built by encoder warps inside the fused forward kernels) and the one with the stand-alone encode kernel
ENC_ALL = nb.FLAG_FUSED_ENCODE_TRAIN  # encoder warps in the training forward too (rendering has them by default)
SCHEDULES = [("bf16", 0), ("fp32_tc", 0), ("bf16", ENC_ALL), ("fp32_tc", ENC_ALL), ("bf16", nb.FLAG_NO_FUSED_ENCODE), ("fp32_tc", nb.FLAG_NO_FUSED_ENCODE),
             ("fp32_tc", nb.FLAG_NO_WEIGHT_MULTICAST), ("bf16", nb.FLAG_NO_WEIGHT_MULTICAST), ("fp32_tc", nb.FLAG_NO_FP8_CORRECTIONS)]
SCHED_IDS = ["bf16", "fp32_tc", "bf16-encoder-warps", "fp32_tc-encoder-warps", "bf16-encode-kernel", "fp32_tc-encode-kernel", "fp32_tc-no-multicast", "bf16-no-multicast",
             "fp32_tc-bf16x3-render"]


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("precision,flags", SCHEDULES, ids=SCHED_IDS)
@pytest.mark.parametrize("R", [200, 1], ids=["R200-chunked", "R1"])
def test_fused_forward_render(R, precision, flags, net):
    """Rendering runs the whole MLP as ONE kernel with TMEM-resident activations (mlp_fused.cu / mlp_fused_split.cu).
    It must agree with the fp64 oracle within the mode's tolerance and with the layer-by-layer chain (same operands,
    same roundings between layers) far tighter; ragged last tile (R*S not a multiple of the tile rows) and the tile loop
    included.  Rendering is deterministic (no jitter) whatever cfg.randomized says."""
    m, ncfg, ocfg = _model(64, precision, engine_flags=flags, **NETS[net])
    m2, _, _ = _model(64, precision, engine_flags=nb.FLAG_NO_FUSED_FORWARD, **NETS[net])
    S = ncfg.n_samples
    rays, pix, _ = batch(R, S)
    params = _params_with_biases(ocfg)
    m.set_params(params)
    m2.set_params(params)
    u = np.zeros((2, R, S + 1), np.float32)
    args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    rgb, depth, acc = m.render(*args)
    rgb2, depth2, acc2 = m2.render(*args)
    ocfg.randomized = 0
    o = orc.train_gradient(ocfg, params, rays, pix, u, with_backward=False, prec="f64")
    print(f"{precision}: fused vs layered rgb {np.abs(rgb - rgb2).max():.2e}  fused vs f64: rgb {np.abs(rgb - o['comp_rgb'][1]).max():.2e} "
          f"acc {np.abs(acc - o['acc'][1]).max():.2e}")
    tol = TOL[precision]
    assert np.isfinite(rgb).all()
    np.testing.assert_allclose(rgb, o["comp_rgb"][1], atol=tol)
    np.testing.assert_allclose(acc, o["acc"][1], atol=tol)
    np.testing.assert_allclose(depth, o["depth"][1], atol=tol * float(rays["fars"].max()))
    np.testing.assert_allclose(rgb, rgb2, atol=0.1 * tol)
    np.testing.assert_allclose(acc, acc2, atol=0.1 * tol)


@pytest.mark.parametrize("precision", ["bf16", "fp32_tc"])
@pytest.mark.parametrize("S,R", [(32, 21), (128, 5), (256, 3)], ids=["S32", "S128", "S256"])
def test_fused_kernels_other_sample_counts(precision, S, R):
    """A 128-row tile is 4 rays at S = 32, one ray at S = 128 (the bench) and half a ray at S = 256: the encoder warps' row ->
    (ray, sample) mapping, the render and a whole gradient step against the fp64 oracle and the per-layer kernels."""
    kw = dict(NET, n_samples=S)
    m, ncfg, ocfg = _model(R, precision, **kw)
    m2, _, _ = _model(R, precision, engine_flags=LAYERED | nb.FLAG_NO_FUSED_ENCODE, **kw)
    rays, pix, u = batch(R, S)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    rgb1, _, acc1 = m.render(*rargs)
    rgb2, _, acc2 = m2.render(*rargs)
    ocfg.randomized = 0
    o = orc.train_gradient(ocfg, params, rays, pix, np.zeros((2, R, S + 1), np.float32), with_backward=False, prec="f64")
    tol = TOL[precision]
    np.testing.assert_allclose(rgb1, o["comp_rgb"][1], atol=tol)
    np.testing.assert_allclose(acc1, o["acc"][1], atol=tol)
    np.testing.assert_allclose(rgb1, rgb2, atol=0.1 * tol)
    assert abs(l1 - l2) <= 1e-5 * abs(l2) and rel_err(g1, g2) <= 2e-3


def _gradient_step(m, params, rays, pix, u):
    m.set_params(params)
    m.set_pixels(pix)
    m.set_sampling_uniforms(u)
    m.GetGradient(rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
    return m.get_gradients().copy(), m.get_loss()[1]


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("precision,flags", SCHEDULES, ids=SCHED_IDS)
def test_fused_training_forward_matches_layered(precision, flags, net):
    """The training forward is the same fused kernel with the activation planes and ReLU bit planes written out for the
    backward pass: a whole gradient step through it must equal the step through the layer-by-layer forward."""
    R = 24
    m, ncfg, ocfg = _model(R, precision, engine_flags=flags, **NETS[net])
    m2, _, _ = _model(R, precision, engine_flags=flags | nb.FLAG_NO_FUSED_TRAIN_FORWARD, **NETS[net])
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    print(f"{precision}: loss {l1:.8f} vs {l2:.8f}; grad max-norm err {rel_err(g1, g2):.2e}")
    assert abs(l1 - l2) <= 1e-6 * abs(l2)
    # accumulation order differs between the two forwards, so a few ReLU masks near zero flip (see the whole-step test)
    assert rel_err(g1, g2) <= 1e-3


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("precision,flags", SCHEDULES, ids=SCHED_IDS)
def test_fused_dgrad_chain_matches_layered(precision, flags, net):
    """The backward dgrad chain of the trunk is one kernel (dZ resident in tensor memory between layers, each layer's dZ
    written once for the wgrad GEMMs).  Same operands, the same ReLU bit masks and the same roundings between layers as
    the per-layer dgrad launches, so the parameter gradients agree to accumulation-order noise."""
    R = 24
    m, ncfg, ocfg = _model(R, precision, engine_flags=flags, **NETS[net])
    m2, _, _ = _model(R, precision, engine_flags=flags | nb.FLAG_NO_FUSED_DGRAD, **NETS[net])
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    g1, _ = _gradient_step(m, params, rays, pix, u)
    g2, _ = _gradient_step(m2, params, rays, pix, u)
    print(f"{precision}: fused vs layered dgrad chain: grad max-norm err {rel_err(g1, g2):.2e}")
    assert np.isfinite(g1).all()
    assert rel_err(g1, g2) <= 1e-5


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("precision,flags", SCHEDULES, ids=SCHED_IDS)
def test_fused_kernels_many_tiles_per_cta(precision, flags, net):
    """More row tiles than CTAs (700 rays x 64 samples = 350 tiles / 175 tile pairs on 148 SMs, ragged tail): every CTA of
    the persistent fused kernels walks several tiles, so the mbarrier phase bookkeeping across tiles is exercised.  Render
    and a whole gradient step must match the layer-by-layer kernels."""
    R = 700
    m, ncfg, ocfg = _model(R, precision, engine_flags=flags, **NETS[net])
    m2, _, _ = _model(R, precision, engine_flags=LAYERED, **NETS[net])
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    rgb1, _, acc1 = m.render(*rargs)
    rgb2, _, acc2 = m2.render(*rargs)
    print(f"{precision}: render rgb diff {np.abs(rgb1 - rgb2).max():.2e}, loss {l1:.8f} vs {l2:.8f}, grad err {rel_err(g1, g2):.2e}")
    tol = TOL[precision]
    assert np.isfinite(rgb1).all() and np.isfinite(g1).all()
    np.testing.assert_allclose(rgb1, rgb2, atol=0.1 * tol)
    np.testing.assert_allclose(acc1, acc2, atol=0.1 * tol)
    assert abs(l1 - l2) <= 1e-5 * abs(l2)
    assert rel_err(g1, g2) <= 2e-3


@pytest.mark.parametrize("precision", ["bf16", "fp32_tc"])
@pytest.mark.parametrize("R", [700, 37], ids=["R700-many-tiles", "R37-ragged"])
def test_inkernel_encoding_is_bit_identical_to_the_encode_kernel(precision, R):
    """accelerated_functions.cu:292-317 + 187-221 inside the fused MLP kernels: the encoder warps run the arithmetic of
    encode.cu (same device functions, same order), so a render and a whole gradient step must give the SAME BITS as the
    path that runs k_encode_pos / k_encode_dir first — through the training planes, through the L2 scratch of the render
    path (more tiles than CTAs: both scratch buffers of every CTA are reused), and on a ragged last tile."""
    m, ncfg, ocfg = _model(R, precision, engine_flags=nb.FLAG_FUSED_ENCODE_TRAIN, **NET)
    m2, _, _ = _model(R, precision, engine_flags=nb.FLAG_NO_FUSED_ENCODE, **NET)
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    out1, out2 = m.render(*rargs), m2.render(*rargs)
    assert l1 == l2
    np.testing.assert_array_equal(g1, g2)
    for a, b in zip(out1, out2):
        np.testing.assert_array_equal(a, b)
    before = m.launch_count()
    m.render(*rargs)
    n1 = m.launch_count() - before
    before = m2.launch_count()
    m2.render(*rargs)
    assert m2.launch_count() - before - n1 >= 2 * 2  # no encode launches on the default path (2 kernels x 2 levels per chunk less)


@pytest.mark.parametrize("precision", ["fp32_tc", "bf16"])
@pytest.mark.parametrize("R", [700, 37, 2], ids=["R700-many-tiles", "R37-ragged", "R2-one-tile"])
def test_weight_multicast_clusters_are_bit_identical_to_single_ctas(R, precision):
    """The fused kernels as 2-CTA clusters that share every weight stage by TMA multicast (mlp_fused_split.cu): the
    arithmetic of a tile does not depend on which CTA walks it or on who fetched its weights, so render, loss and gradients
    must be the same bits as with one CTA per slot — with more tiles than CTAs (rings recycled many times), an odd tile count
    (a phantom tile keeps the pair's rings in lock-step) and fewer tiles than one cluster."""
    m, ncfg, ocfg = _model(R, precision, **NET)
    m2, _, _ = _model(R, precision, engine_flags=nb.FLAG_NO_WEIGHT_MULTICAST, **NET)
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    out1, out2 = m.render(*rargs), m2.render(*rargs)
    assert l1 == l2
    np.testing.assert_array_equal(g1, g2)
    for a, b in zip(out1, out2):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("precision,flags", SCHEDULES, ids=SCHED_IDS)
def test_tc_step_is_bitwise_reproducible(precision, flags):
    """No atomics anywhere on the tensor-core path (wgrad partials and head partials are reduced in a fixed order): the same
    batch gives bit-identical gradients, loss and rendered pixels on every run."""
    R = 48
    m, ncfg, ocfg = _model(R, precision, engine_flags=flags, **NET)
    S = ncfg.n_samples
    rays, pix, u = batch(R, S)
    m.set_pixels(pix)
    m.set_sampling_uniforms(u)
    args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"])
    runs = []
    for _ in range(3):
        m.GetGradient(*args)
        runs.append((m.get_gradients().copy(), m.get_loss()[1], m.render(*args[:5])[0].copy()))
    for g, l, img in runs[1:]:
        np.testing.assert_array_equal(g, runs[0][0])
        assert l == runs[0][1]
        np.testing.assert_array_equal(img, runs[0][2])


# ---------------------------------------------------------------------------------------------- NERF_FLAG_WGRAD_FP16
W16 = nb.FLAG_WGRAD_FP16


def _per_tensor(m, a, b):
    out, off = [], 0
    for n in m.GetLayerSizes():
        out.append(rel_err(a[off:off + n], b[off:off + n]))
        off += n
    return out


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("S,R", [(64, 64), (128, 40), (256, 9), (32, 70)], ids=["S64", "S128", "S256", "S32"])
def test_wgrad_fp16_option_against_the_three_term_wgrad(net, S, R):
    """NERF_FLAG_WGRAD_FP16 (opt-in, fp32-accurate mode): activations / encodings / dZ leave the fused kernels as ONE fp16 plane
    each and wgrad multiplies them with one fp16 MMA per product.  The forward arithmetic is untouched — loss and rendered
    pixels are the SAME BITS as without the flag — and on a real step (every gradient element a sum over thousands of samples)
    the parameter gradient stays within the mode's 1e-4 of the default three-term wgrad and of the fp64 oracle."""
    kw = dict(NETS[net], n_samples=S)
    bf16x3 = nb.FLAG_NO_FP8_CORRECTIONS  # the same forward kernels on both sides (the option alone also switches the forward products)
    m, ncfg, ocfg = _model(R, "fp32_tc", engine_flags=W16 | bf16x3, **kw)
    m2, _, _ = _model(R, "fp32_tc", engine_flags=bf16x3, **kw)
    rays, pix, u = batch(R, S)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    assert l1 == l2
    assert np.isfinite(g1).all()
    o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64")
    per = _per_tensor(m, g1, g2)
    print(f"{net} S={S} R={R}: fp16-wgrad vs three-term wgrad: whole {rel_err(g1, g2):.2e}, worst tensor {max(per):.2e} (#{int(np.argmax(per))}); "
          f"vs fp64 {rel_err(g1, o64['grads']):.2e} (three-term: {rel_err(g2, o64['grads']):.2e})")
    assert rel_err(g1, g2) <= 1e-4
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    for a, b in zip(m.render(*rargs), m2.render(*rargs)):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("R", [700, 37], ids=["R700-many-tiles", "R37-ragged"])
def test_wgrad_fp16_option_is_reproducible_and_cluster_independent(R):
    """The fp16 planes go through the same per-warp TMA boxes: bit-identical between 2-CTA clusters and single CTAs, between
    runs, with more tiles than CTAs and on a ragged last tile."""
    bf16x3 = nb.FLAG_NO_FP8_CORRECTIONS
    m, ncfg, ocfg = _model(R, "fp32_tc", engine_flags=W16 | bf16x3, **NET)
    m2, _, _ = _model(R, "fp32_tc", engine_flags=W16 | bf16x3 | nb.FLAG_NO_WEIGHT_MULTICAST, **NET)
    m3, _, _ = _model(R, "fp32_tc", engine_flags=bf16x3, **NET)
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    g1b, l1b = _gradient_step(m, params, rays, pix, u)
    assert l1 == l2 == l1b
    np.testing.assert_array_equal(g1, g2)
    np.testing.assert_array_equal(g1, g1b)
    g3, _ = _gradient_step(m3, params, rays, pix, u)
    assert rel_err(g1, g3) <= 1e-4


def test_wgrad_fp16_option_per_gemm_bound_on_random_walk_sums():
    """The honest limit of the option, and why it is not the default: with synthetic, sign-random upstream gradients every
    dW element is a random walk, nothing averages out, and the 2^-12 rounding of the two fp16 operands stays visible at
    ~3e-4 .. 1e-3 of the tensor scale (CPU simulation: scripts/wgrad_fp16_precision.py) — the default
    three-term wgrad passes the same inputs at 1e-4 (test_tc_mlp_backward)."""
    R = 64
    m, ncfg, ocfg = _model(R, "fp32_tc", engine_flags=W16, **NET)
    S = ncfg.n_samples
    M = R * S
    rng = np.random.default_rng(4)
    P, Dd = 6 * ncfg.deg_point, 3 + 6 * ncfg.deg_view
    params = _params_with_biases(ocfg)
    m.set_params(params)
    ep, ed = _stable_inputs(ocfg, params, M, P, Dd, 2e-4, 4)
    m.mlp.get_output(dev(ep), dev(ed), 1, R)
    cg, dg = rng.normal(size=(M, 3)).astype(np.float32), rng.normal(size=M).astype(np.float32)
    m.mlp.reset_gradients(1)
    m.mlp.get_gradient(dev(cg), dev(dg), 1)
    rd64, rr64, acts64 = orc.mlp_forward(ocfg, params, ep, ed, prec="f64")
    d_rd, d_rr = orc.output_activations_grad(ocfg, rd64, rr64, dg, cg, prec="f64")
    g64 = orc.mlp_backward(ocfg, params, ep, ed, acts64, d_rd, d_rr, prec="f64")
    per = _per_tensor(m, m.get_gradients().astype(np.float64), g64)
    print(f"fp16-wgrad option, random-walk sums, M={M}: per-tensor max-norm err " + " ".join(f"{e:.1e}" for e in per))
    assert max(per) <= 2e-3
    # the per-level scale of the fp16 dZ planes is a power of two taken from the head gradients: upstream gradients 2^-20 or
    # 2^+12 times as large give the same fp16 planes and EXACTLY the scaled parameter gradient
    g_ref = m.get_gradients().copy()
    for k in (-20, 12):
        f = np.float32(2.0 ** k)
        m.mlp.reset_gradients(1)
        m.mlp.get_gradient(dev(cg * f), dev(dg * f), 1)
        np.testing.assert_array_equal(m.get_gradients(), g_ref * f)


def test_wgrad_fp16_option_needs_the_fused_fp32_accurate_path():
    for kw in (dict(precision="bf16"), dict(precision="fp32_tc", engine_flags_extra=nb.FLAG_NO_FUSED_DGRAD), dict(precision="fp32")):
        prec = kw["precision"]
        flags = W16 | kw.get("engine_flags_extra", 0)
        if prec == "fp32":
            nb.AcceleratedMipNeRF(nb.default_config(n_rays=8, precision="fp32", engine_flags=flags, **NET)).close()  # CUDA-core engine: flags ignored
            continue
        with pytest.raises(nb.NerfError):
            nb.AcceleratedMipNeRF(nb.default_config(n_rays=8, precision=prec, engine_flags=flags, **NET))


# ---------------------------------------------------------------------------------------------- fp16 + fp8-correction products
NO_F8C = nb.FLAG_NO_FP8_CORRECTIONS


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("flags", [0, nb.FLAG_NO_WEIGHT_MULTICAST, nb.FLAG_NO_FUSED_ENCODE], ids=["clusters", "single-ctas", "encode-kernel"])
@pytest.mark.parametrize("R", [200, 700, 1], ids=["R200", "R700-many-tiles", "R1"])
def test_fp8_corrections_render(R, flags, net):
    """Rendering in the fp32-accurate mode computes a w as fp16 x fp16 plus two E4M3 correction products onto the same accumulator
    (mlp_fused_split.cu, REP = 1; NERF_FLAG_NO_FP8_CORRECTIONS keeps the three bf16 products).  Rendered pixels must stay within
    the mode's 1e-4 of the fp64 oracle — measured: as close as the bf16x3 kernels — and close to those kernels; ragged tiles,
    more tiles than CTAs, clusters or single CTAs, encoder warps or the encode kernel."""
    m, ncfg, ocfg = _model(64, "fp32_tc", engine_flags=flags, **NETS[net])
    m2, _, _ = _model(64, "fp32_tc", engine_flags=NO_F8C, **NETS[net])
    S = ncfg.n_samples
    rays, pix, _ = batch(R, S)
    params = _params_with_biases(ocfg)
    m.set_params(params)
    m2.set_params(params)
    args = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    rgb, depth, acc = m.render(*args)
    rgb2, depth2, acc2 = m2.render(*args)
    ocfg.randomized = 0
    o = orc.train_gradient(ocfg, params, rays, pix, np.zeros((2, R, S + 1), np.float32), with_backward=False, prec="f64")
    print(f"{net} R={R}: fp8-corrections vs f64: rgb {np.abs(rgb - o['comp_rgb'][1]).max():.2e} acc {np.abs(acc - o['acc'][1]).max():.2e}; "
          f"bf16x3 vs f64: rgb {np.abs(rgb2 - o['comp_rgb'][1]).max():.2e}; fp8-corrections vs bf16x3 {np.abs(rgb - rgb2).max():.2e}")
    assert np.isfinite(rgb).all()
    np.testing.assert_allclose(rgb, o["comp_rgb"][1], atol=1e-4)
    np.testing.assert_allclose(acc, o["acc"][1], atol=1e-4)
    np.testing.assert_allclose(depth, o["depth"][1], atol=1e-4 * float(rays["fars"].max()))
    np.testing.assert_allclose(rgb, rgb2, atol=5e-5)
    again = m.render(*args)
    np.testing.assert_array_equal(again[0], rgb)


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("S,R", [(64, 64), (128, 40)], ids=["S64", "S128"])
def test_fp8_corrections_training_step(net, S, R):
    """Training with NERF_FLAG_WGRAD_FP16 runs the fp8-correction forward too: it writes fp16(32 a) planes (the wgrad GEMMs divide
    the factor out).  Against the same option on the bf16x3 forward the loss agrees to fp32 noise and the gradient to the few
    ReLU masks that differ between two forward arithmetics (tests/test_bench_config_parity_gpu.py measures them: 4.8e-6 against
    fp64, gradient on the same branches 7.6e-6)."""
    kw = dict(NETS[net], n_samples=S)
    m, ncfg, ocfg = _model(R, "fp32_tc", engine_flags=W16, **kw)
    m2, _, _ = _model(R, "fp32_tc", engine_flags=W16 | NO_F8C, **kw)
    rays, pix, u = batch(R, S)
    params = _params_with_biases(ocfg)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    g1b, l1b = _gradient_step(m, params, rays, pix, u)
    o64 = orc.train_gradient(ocfg, params, rays, pix, u, prec="f64")
    per = _per_tensor(m, g1, g2)
    print(f"{net} S={S}: fp8-corrections + fp16-wgrad vs default: loss {l1:.8f} / {l2:.8f} (f64 {o64['total_loss']:.8f}), gradient whole {rel_err(g1, g2):.2e}, "
          f"worst tensor {max(per):.2e}; vs fp64 {rel_err(g1, o64['grads']):.2e} (default {rel_err(g2, o64['grads']):.2e})")
    assert l1 == l1b
    np.testing.assert_array_equal(g1, g1b)
    assert abs(l1 - o64["total_loss"]) <= 1e-4 * o64["total_loss"]
    assert np.isfinite(g1).all() and rel_err(g1, g2) <= 1e-3  # same bound as fused vs per-layer forward (test_fused_training_forward_matches_layered)


# ---------------------------------------------------------------------------------------------- NERF_FLAG_PAIR_MMA
@pytest.mark.parametrize("flags", [0, nb.FLAG_NO_FP8_CORRECTIONS, nb.FLAG_WGRAD_FP16], ids=["default", "bf16x3-render", "wgrad-fp16"])
@pytest.mark.parametrize("R", [700, 37, 2], ids=["R700-many-tiles", "R37-ragged", "R2-one-tile"])
def test_pair_mma_is_bit_identical_to_multicast_clusters(R, flags):
    """NERF_FLAG_PAIR_MMA: the 2-CTA clusters of the fp32-accurate fused kernels issue tcgen05.mma.cta_group::2 — one MMA for both
    tiles, each SM holding only its half of every weight tile — instead of sharing whole tiles by multicast.  Every output row is
    the same sequence of products, so render, loss and gradients are the same bits; more tiles than CTAs, an odd tile count
    (phantom tile), fewer tiles than a cluster."""
    m, ncfg, ocfg = _model(R, "fp32_tc", engine_flags=flags | nb.FLAG_PAIR_MMA, **NET)
    m2, _, _ = _model(R, "fp32_tc", engine_flags=flags, **NET)
    rays, pix, u = batch(R, ncfg.n_samples)
    params = _params_with_biases(ocfg)
    rargs = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"])
    out1, out2 = m.render(*rargs), m2.render(*rargs)
    for a, b in zip(out1, out2):
        np.testing.assert_array_equal(a, b)
    g1, l1 = _gradient_step(m, params, rays, pix, u)
    g2, l2 = _gradient_step(m2, params, rays, pix, u)
    assert l1 == l2
    np.testing.assert_array_equal(g1, g2)
