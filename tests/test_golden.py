"""Committed golden vector (tests/golden/step_small.npz, made by tests/golden/make_golden.py): the oracle on CPU and
the CUDA path on the GPU must both keep reproducing it."""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden.make_golden import CFG, R

G = Path(__file__).parent / "golden" / "step_small.npz"


def _load():
    z = np.load(G)
    rays = {k[4:]: z[k] for k in z.files if k.startswith("ray_")}
    return z, rays


def _scale_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())


def test_oracle_reproduces_golden():
    z, rays = _load()
    cfg = orc.default_config(**CFG)
    np.testing.assert_array_equal(orc.init_params(cfg, 7), z["params"])  # Philox init stream is frozen too
    for prec, tol in (("f64", 2e-7), ("f32", 1e-4)):  # the fixture is stored as float32
        o = orc.train_gradient(cfg, z["params"], rays, z["pixels"], z["u"], prec=prec)
        assert _scale_err(o["grads"], z["grads"]) <= tol
        assert _scale_err(o["comp_rgb"], z["comp_rgb"]) <= tol
        assert _scale_err(o["weights"], z["weights"]) <= 10 * tol
        assert abs(o["total_loss"] - float(z["total_loss"])) <= tol * float(z["total_loss"]) + 1e-7
    np.testing.assert_array_equal(orc.train_gradient(cfg, z["params"], rays, z["pixels"], z["u"], prec="f32")["t_vals"][0], z["t_vals"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp32_tc", 1e-3), ("bf16", 5e-2)])
def test_cuda_path_reproduces_golden(precision, tol):
    import nerf_or_nothing_b200 as nb

    z, rays = _load()
    m = nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, precision=precision, **CFG))
    opt = nb.AcceleratedAdamOptimizer(m.GetLayerSizes())
    m.set_params(z["params"])
    m.set_sampling_uniforms(z["u"])
    loss = m.train_step(opt, rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                        rays["loss_mults"], z["pixels"], 1e-3)
    out_tol = {"fp32": 1e-4, "fp32_tc": 1e-4, "bf16": 2e-2}[precision]
    assert abs(loss - float(z["total_loss"])) <= out_tol * float(z["total_loss"])
    assert _scale_err(m.get_gradients(), z["grads"]) <= tol  # whole-step gradients: see DESIGN.md §4 (ReLU kinks)
    from tests.gpu_util import from_ptr

    S = CFG["n_samples"]
    for lv in range(2):
        o = m.level_outputs(lv)
        assert _scale_err(from_ptr(o["comp_rgb"], (R, 3)), z["comp_rgb"][lv]) <= out_tol
        assert _scale_err(from_ptr(o["acc"], (R,)), z["acc"][lv]) <= out_tol
    np.testing.assert_array_equal(from_ptr(m.level_outputs(0)["t_vals"], (R, S + 1)), z["t_vals"][0])
    if precision == "fp32":
        assert np.abs(m.get_params() - z["params_after_adam"]).max() <= 2e-5
