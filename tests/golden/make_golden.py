"""Regenerates tests/golden/step_small.npz — a frozen whole-step input/output vector.

The reference ships no golden vectors and its CPU path is C# (not runnable here), so this fixture is
produced by the fp64 build of the CPU oracle (oracle/oracle.c, pinned by tests/test_oracle_cpu.py against torch-fp64
autograd and, on the GPU box, against the reference's own CUDA kernels).  It freezes today's semantics: any later
change to the oracle OR the CUDA path that moves these numbers is caught on both sides.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402

CFG = dict(n_samples=32, net_depth=4, net_width=64, net_depth_condition=1, net_width_condition=64, skip_layer=2,
           deg_point=8, deg_view=2, density_bias=-1.0, rgb_padding=0.001)
R = 48


def inputs():
    rays, pix = orc.synthetic_rays(R, width=100, height=100, seed=4242)
    rays["loss_mults"] = np.random.default_rng(9).uniform(0.5, 1.5, R).astype(np.float32)
    u = np.stack([orc.sampling_uniforms(99, 3, lv, 0, R, CFG["n_samples"] + 1) for lv in range(2)])
    return rays, pix, u


def main():
    cfg = orc.default_config(**CFG)
    rays, pix, u = inputs()
    params = orc.init_params(cfg, 7)
    o = orc.train_gradient(cfg, params, rays, pix, u, prec="f64")
    p1, m1, v1 = orc.adam_step(params, o["grads"], np.zeros_like(o["grads"]), np.zeros_like(o["grads"]), 1e-3, 1, 0, prec="f64")
    out = Path(__file__).with_name("step_small.npz")
    np.savez_compressed(out, params=params, u=u, pixels=pix, **{"ray_" + k: v for k, v in rays.items()},
                        grads=o["grads"].astype(np.float32), comp_rgb=o["comp_rgb"].astype(np.float32),
                        depth=o["depth"].astype(np.float32), acc=o["acc"].astype(np.float32),
                        t_vals=o["t_vals"].astype(np.float32), weights=o["weights"].astype(np.float32),
                        loss=o["loss"].astype(np.float32), total_loss=np.float32(o["total_loss"]),
                        params_after_adam=p1.astype(np.float32))
    print(out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
