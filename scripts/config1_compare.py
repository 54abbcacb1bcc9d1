#!/usr/bin/env python
"""BASELINE.json configs[0] (SURVEY §8d "C1"): synthetic 100x100 Blender-style scene, 1024-ray batch, coarse+fine MipNeRF MLP
8x256, 64+64 samples, 100 train steps — on the CPU restatement of the reference's C# path (oracle/, all host cores) AND on
one GPU, with identical batches, sampling uniforms, initial weights and learning rates.  Reports the two loss curves,
their maximum relative deviation, the final-parameter difference and the two step times.

    python scripts/config1_compare.py [--steps 100] [--out gpurun_out/config1_compare.json]

This is a checker script (it may use the oracle, like tests/): nothing in the product imports it.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import nerf_or_nothing_b200 as nb  # noqa: E402
from nerf_or_nothing_b200.scene import synthetic_rays  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "config1_compare.json"))
    ap.add_argument("--precisions", default="fp32,fp32_tc,bf16")
    ap.add_argument("--window", type=int, default=50, help="steps per window of the windowed loss deviation")
    a = ap.parse_args()
    R, S = a.rays, 64
    ocfg = orc.default_config(n_samples=S)
    models = {p: nb.AcceleratedMipNeRF(nb.default_config(n_rays=R, n_samples=S, precision=p)) for p in a.precisions.split(",")}
    opts = {p: nb.AcceleratedAdamOptimizer(m.GetLayerSizes()) for p, m in models.items()}
    params = orc.init_params(ocfg, 7)
    for m in models.values():
        m.set_params(params)
    mo, vo = np.zeros_like(params), np.zeros_like(params)
    curves = {k: [] for k in ("cpu", *models)}
    t_cpu, t_gpu = [], {p: [] for p in models}
    for step in range(1, a.steps + 1):
        rays, pix = synthetic_rays(R, width=100, height=100, n_views=100, seed=2024 + step)
        u = np.stack([orc.sampling_uniforms(99, step, lv, 0, R, S + 1) for lv in range(2)])
        lr = 5e-4
        for p, m in models.items():
            m.set_sampling_uniforms(u)
            t0 = time.perf_counter()
            curves[p].append(m.train_step(opts[p], rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"],
                                          rays["loss_mults"], pix, lr))
            t_gpu[p].append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        o = orc.train_gradient(ocfg, params, rays, pix, u, prec="f32")
        params, mo, vo = orc.adam_step(params, o["grads"], mo, vo, lr, step, 0, prec="f32")
        t_cpu.append(time.perf_counter() - t0)
        curves["cpu"].append(o["total_loss"])
        if step % 10 == 0 or step == 1:
            print(f"step {step}: cpu {curves['cpu'][-1]:.6f} " + " ".join(f"{p} {curves[p][-1]:.6f}" for p in models), flush=True)
    cpu = np.asarray(curves["cpu"])
    out = {"config": f"configs[0]: 100x100 scene, {R}-ray batch, 8x256 MLP, {S}+{S} samples, {a.steps} steps", "cpu_threads": orc.max_threads(),
           "cpu_median_s_per_step": float(np.median(t_cpu)), "cpu_rays_per_s": R / float(np.median(t_cpu)), "loss_cpu": curves["cpu"]}
    for p, m in models.items():
        c = np.asarray(curves[p])
        nw = len(c) // a.window
        wc, wr = c[:nw * a.window].reshape(nw, a.window).mean(1), cpu[:nw * a.window].reshape(nw, a.window).mean(1)
        out[p] = {"loss": curves[p], "max_rel_loss_deviation_vs_cpu": float(np.max(np.abs(c - cpu) / cpu)),
                  "window": a.window, "windowed_rel_deviation_vs_cpu": [float(x) for x in np.abs(wc - wr) / wr],
                  "max_windowed_rel_deviation_vs_cpu": float(np.max(np.abs(wc - wr) / wr)),
                  "max_windowed_deviation_over_initial_loss": float(np.max(np.abs(wc - wr)) / wr[0]),
                  "final_param_max_abs_diff_vs_cpu": float(np.abs(m.get_params() - params).max()),
                  "final_param_rel_l2_diff_vs_cpu": float(np.linalg.norm(m.get_params() - params) / np.linalg.norm(params)),
                  "gpu_median_ms_per_step_e2e": 1e3 * float(np.median(t_gpu[p])), "gpu_rays_per_s_e2e": R / float(np.median(t_gpu[p]))}
        print(p, {k: v for k, v in out[p].items() if k != "loss"})
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(out))
    print("cpu", out["cpu_median_s_per_step"], "s/step on", out["cpu_threads"], "threads ->", a.out)


if __name__ == "__main__":
    main()
