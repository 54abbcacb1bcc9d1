#!/usr/bin/env python
"""The reference's OWN CUDA kernels (ANU/accelerated_functions.cu, compiled unmodified for sm_100a into
oracle/_ref/libref_kernels.so) timed on the B200 beside the kernels of libnerfb200.so that replace them, stage by stage,
at the reference's compile-time problem size (1024 rays x 128 samples, .cu:15-16) — BASELINE.md §4 "GPU-side reference
baseline", SURVEY §2.2 "the on-box baseline is the reference kernel itself recompiled".

    python scripts/time_reference_kernels.py [--out gpurun_out/ref_kernel_times.json]

Checker-side script (it loads oracle/_ref like tests/): nothing in the product imports it.  Times are CUDA events on the
launching (default) stream, best of `reps` after a warm-up; both sides' entry points synchronise after the launch, so an
event pair brackets exactly one kernel.  Clocks are sampled during the run (bench.ClockSampler).
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import nerf_or_nothing_b200 as nb  # noqa: E402
from oracle import oracle as orc  # noqa: E402

R, S = 1024, 128
M = R * S


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def timed(fn, reps, warm=True):
    if warm:
        fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "ref_kernel_times.json"))
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    ref, lib = orc.ref_lib(), nb.lib()
    assert ref.ref_num_rays() == R and ref.ref_num_samples() == S

    def rc(name, *args):
        r = getattr(ref, name)(*[C.c_float(x) if isinstance(x, float) else x for x in args])
        assert r == 0, (name, r)

    def nc(name, *args):
        nb.check(getattr(lib, name)(*args))

    clocks = bench.ClockSampler(0)
    clocks.start()
    time.sleep(0.25)
    t_begin = time.time()
    rng = np.random.default_rng(0)
    rays, pix = orc.synthetic_rays(R, width=800, height=800)
    t = orc.sample_t_vals(rays["nears"], rays["fars"], orc.sampling_uniforms(99, 0, 0, 0, R, S + 1), S)
    dt, do, dd, dr = dev(t), dev(rays["origins"]), dev(rays["directions"]), dev(rays["radii"])
    rows = []

    def row(stage, ref_kernel, new_kernel, f_ref, f_new, note="", ref_reps=None):
        tr = timed(f_ref, ref_reps or a.reps, warm=ref_reps is None)
        tn = timed(f_new, a.reps)
        rows.append({"stage": stage, "reference_kernel": ref_kernel, "reference_ms": round(tr, 4), "new_kernel": new_kernel,
                     "new_ms": round(tn, 4), "speedup": round(tr / tn, 1), "note": note})
        print(f"{stage:34s} reference {tr:10.3f} ms   new {tn:8.3f} ms   x{tr / tn:8.1f}   {note}", flush=True)

    mean, cov = torch.empty(M, 3, device="cuda"), torch.empty(M, 3, device="cuda")
    row("cast_rays", "cast_rays (.cu:292-317)", "k_cast_rays",
        lambda: rc("ref_cast_rays", ptr(dt), ptr(do), ptr(dd), ptr(mean), ptr(cov), ptr(dr)),
        lambda: nc("nerf_cast_rays", ptr(dt), ptr(do), ptr(dd), ptr(mean), ptr(cov), ptr(dr), R, S))
    dps = dev(np.repeat(rays["directions"], S, axis=0))
    ep, ed = torch.empty(M, 96, device="cuda"), torch.zeros(M, 27, device="cuda")
    row("encode_input_data", "encode_input_data (.cu:187-221)", "k_encode_pos<false> + k_encode_dir",
        lambda: rc("ref_encode_input_data", ptr(mean), ptr(cov), ptr(dps), ptr(ep), ptr(ed)),
        lambda: nc("nerf_encode_input_data", ptr(mean), ptr(cov), ptr(dd), ptr(ep), ptr(ed), R, S, 16, 4),
        "fp32 [M,96] + [M,27] out; the training/render paths build the encodings inside the fused MLP kernels instead")
    # dense layers, 256 x 256 (the trunk shape): forward, then backward (the reference: 2 global float atomics per multiply)
    n = k = 256
    x = dev(rng.uniform(0, 1, (M, k)))
    W = dev(rng.normal(size=(n, k)) / np.sqrt(k))
    b = dev(rng.normal(size=n) * 0.1)
    y, z = torch.empty(M, n, device="cuda"), torch.empty(M, n, device="cuda")
    row("dense layer fwd 256x256 (fp32)", "get_neuron_output (.cu:36-48)", "k_sgemm<1,1> (per-stage fp32 entry)",
        lambda: rc("ref_apply_layer", 0, ptr(x), ptr(W), ptr(b), ptr(y), ptr(z), n, k),
        lambda: nc("nerf_apply_layer", ptr(x), None, ptr(W), ptr(b), ptr(y), ptr(z), M, n, k, 0, 0),
        "the per-stage entry is the strict-fp32 CUDA-core kernel; the model runs the fused tcgen05 kernels")
    dy = dev(rng.normal(size=(M, n)) * 1e-3)
    gi, gW, gb = torch.zeros(M, k, device="cuda"), torch.zeros(n, k, device="cuda"), torch.zeros(n, device="cuda")
    row("dense layer bwd 256x256 (fp32)", "backpropagate_neuron (.cu:91-111)", "k_sgemm<1,0> + k_sgemm<0,0> + reduction",
        lambda: rc("ref_backpropagate_layer", 0, ptr(x), ptr(W), ptr(z), ptr(dy), ptr(gi), ptr(gW), ptr(gb), n, k),
        lambda: nc("nerf_backpropagate_layer", ptr(x), None, ptr(W), ptr(z), ptr(dy), ptr(gi), ptr(gW), ptr(gb), M, n, k, 0, 0),
        "reference: one launch, no warm-up (2 global float atomics per multiply: 17 G atomics)", ref_reps=1)
    rgb = dev(rng.uniform(0, 1, (R, S, 3)))
    den = dev(rng.uniform(0, 1, (R, S)) ** 4 * 30)
    comp, al, tr_, w = (torch.empty(R, 3, device="cuda"), torch.empty(R, S, device="cuda"), torch.empty(R, S, device="cuda"),
                        torch.empty(R, S, device="cuda"))
    row("volumetric_rendering", "volumetric_rendering (.cu:318-344)", "k_composite_fwd<16,false>",
        lambda: rc("ref_volumetric_rendering", ptr(rgb), ptr(den), ptr(dt), ptr(dd), ptr(comp), ptr(al), ptr(tr_), ptr(w)),
        lambda: nc("nerf_volumetric_rendering", ptr(rgb), ptr(den), ptr(dt), ptr(dd), ptr(comp), None, None, ptr(w), R, S, 1))
    g = dev(rng.normal(size=(R, 3)))
    grgb, gden = torch.zeros(R, S, 3, device="cuda"), torch.zeros(R, S, device="cuda")
    row("volumetric_rendering_gradient", "volumetric_rendering_gradient (.cu:362-402)", "k_composite_bwd<16,false>",
        lambda: rc("ref_volumetric_rendering_gradient", ptr(g), ptr(al), ptr(tr_), ptr(w), ptr(rgb), ptr(dt), ptr(dd), ptr(grgb), ptr(gden)),
        lambda: nc("nerf_volumetric_rendering_gradient", ptr(g), ptr(rgb), ptr(den), ptr(dt), ptr(dd), ptr(grgb), ptr(gden), R, S, 1, 1))
    lm = dev(rng.uniform(0.5, 2, R))
    pixd, gg = dev(pix), torch.zeros(R, 3, device="cuda")
    lms = float(lm.sum().item())
    row("get_output_gradient", "get_output_gradient (.cu:347-361)", "k_output_gradient",
        lambda: rc("ref_get_output_gradient", ptr(comp), ptr(pixd), ptr(lm), ptr(gg), lms, 0),
        lambda: nc("nerf_get_output_gradient", ptr(comp), ptr(pixd), ptr(lm), ptr(gg), lms, 0.1, R))
    P = 546948
    p_, g_, m_, v_ = dev(rng.normal(size=P)), dev(rng.normal(size=P) * 1e-3), torch.zeros(P, device="cuda"), torch.zeros(P, device="cuda")
    # the reference launches Adam once per tensor (22 launches, ANU/AcceleratedAdamOptimizer.cpp:31-39); one flat call here
    row("adam_optimizer_step (546948 params)", "adam_optimizer_step (.cu:403-416), one launch", "k_adam (one float4 pass)",
        lambda: rc("ref_adam_optimizer_step", ptr(p_), ptr(g_), ptr(m_), ptr(v_), 1e-3, 0.9, 0.999, 1.0, 1.0, P),
        lambda: nc("nerf_adam_optimizer_step", ptr(p_), ptr(g_), ptr(m_), ptr(v_), 1e-3, 0.9, 0.999, 1.0, 1.0, P, 0))
    clk = clocks.stop(t_begin, time.time())
    out = {"what": "reference kernels (accelerated_functions.cu recompiled for sm_100a, unmodified) vs libnerfb200 per-stage entries, "
                   f"{R} rays x {S} samples, best of {a.reps}, CUDA events around one synchronous launch each",
           "gpu": torch.cuda.get_device_name(0), "clocks": clk, "stages": rows}
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(out, indent=1))
    print("->", a.out)


if __name__ == "__main__":
    main()
