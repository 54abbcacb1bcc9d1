#!/usr/bin/env python
"""bench.py — train rays/sec (fwd + bwd + Adam) of the MipNeRF hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|fp32_tc|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 800x800 Blender-shape scene, 4096-ray batch PER GPU, 128+128 samples,
8x256 MipNeRF MLP (546 948 params), fp32 training.  One "step" = one pass of the hot path over one ray batch:
sample t -> cast_rays+IPE -> MLP fwd -> compositing (both levels) -> MSE gradient -> compositing bwd -> MLP bwd ->
[NCCL allreduce of the 2.19 MB gradient when N > 1] -> Adam.  Rays shard across ranks (weak scaling: 4096 rays/GPU;
N=8 is the 32 768-ray global batch of configs[2]).

`value`  device-resident inputs, timed with CUDA events on the library's stream, max over ranks.
`e2e`    the same steps through the reference-facing call with HOST arrays (H2D inside, loss read back each step).
`roofline` the dominant kernel family, timed live inside the timed region with in-stream CUDA events.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference's C# path (oracle/, kind "port" — the C#
cannot be built here), all host threads, on a bounded ray sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RAYS_PER_GPU = 4096
N_SAMPLES = 128
METRIC = "train rays/sec (fwd+bwd+Adam)"
UNIT = "rays/s"


def model_kw():
    return dict(n_samples=N_SAMPLES, net_depth=8, net_width=256, net_depth_condition=1, net_width_condition=128,
                skip_layer=4, deg_point=16, deg_view=4)


def layer_table():
    D, W, Wc, P, Dd = 8, 256, 128, 96, 27
    layers = [(W, P, 0)] + [(W, W, P if i % 4 == 0 else 0) for i in range(1, D)] + [(1, W, 0), (Wc, W, Dd), (3, Wc, 0)]
    return layers


W16_MODE = "fp32_tc+wgrad_fp16"  # the fp32-accurate mode with NERF_FLAG_WGRAD_FP16 (opt-in): sub-record `modes[W16_MODE]`
W16_FLAG = 128


def gemm_bytes(R, S, precision, levels=2):
    """Algorithmic HBM bytes per step of the GEMM families and the thin heads: the MINIMUM any schedule must
    move given what later stages read — every stored activation / gradient element crosses HBM once per use, at the
    element size of the mode (fp32-accurate mode: 4 B = hi + lo bf16 planes; bf16 mode: 2 B).
    forward: reads the encodings, writes each layer's activations + ReLU bits (the backward pass needs them);
    dgrad:   reads dZ of the condition layer and the ReLU bits, writes dZ of every trunk layer (wgrad needs them) — a
             per-layer implementation additionally re-reads each dZ it just wrote (2x), which this figure does not credit;
    wgrad:   reads dZ and X of every layer;   heads backward: reads X_cond, X_trunk, writes dZ_cond."""
    M = R * S * levels
    eb = 2 if precision == "bf16" else 4
    # planes that exist only for wgrad (activations, dZ): one fp16 plane in the W16 mode; the encodings and dZ of the
    # condition layer are ALSO operands of the fused forward / dgrad chain and stay hi + lo there
    sb = 2 if precision == W16_MODE else eb
    dense = [l for l in layer_table() if l[0] > 4]
    fwd = M * eb * (128 + 64) + sum(M * (sb * o + o // 8) for o, a, b in dense)
    dgrad = M * eb * dense[-1][0] + sum(M * (sb * a + a // 8) for i, (o, a, b) in enumerate(dense) if i > 0)
    wgrad = sum(M * sb * (o + a + b) for o, a, b in dense)
    heads = M * (sb * (256 + 128) + (eb + (sb if sb != eb else 0)) * 128 + 128 // 8 + 16)
    return {"mlp_fwd_gemm": fwd, "mlp_dgrad_gemm": dgrad, "mlp_wgrad_gemm": wgrad, "mlp_bwd_heads": heads}


def algorithmic_work(R, S, levels=2):
    """Per-step algorithmic FLOPs / bytes of each kernel family."""
    M = R * S * levels
    lay = layer_table()
    dense = [l for l in lay if l[0] > 4]
    fwd = 2 * M * sum(o * (a + b) for o, a, b in dense)
    wgrad = fwd
    dgrad = 2 * M * sum(o * a for i, (o, a, b) in enumerate(dense) if i > 0)  # no gradient into the encodings
    n_params = sum(o * (a + b) + o for o, a, b in lay)
    return {
        "mlp_fwd_gemm": ("TFLOP/s", fwd), "mlp_wgrad_gemm": ("TFLOP/s", wgrad), "mlp_dgrad_gemm": ("TFLOP/s", dgrad),
        "mlp_fwd_heads": ("GB/s", M * 4 * (256 + 128 + 4)),  # fp32 CUDA-core mode only (the TC modes fuse the heads into the GEMM epilogue)
        "composite_fwd": ("GB/s", M * 24 + levels * R * 32), "composite_bwd": ("GB/s", M * 36 + levels * R * 24),
        "cast_rays+encode": ("GB/s", M * (4 + 96 * 4 + 28 * 4) + levels * R * 32),
        "adam": ("GB/s", n_params * 28), "sample_t_vals": ("GB/s", M * 8), "loss_gradient": ("GB/s", levels * R * 40),
    }, n_params


def config_dict(R, world, precision, global_batch=0):
    n_params = sum(o * (a + b) + o for o, a, b in layer_table())
    if global_batch:
        what = (f"configs[2]: same 800x800 Blender-shape scene, {global_batch}-ray global batch split over {world} GPU(s) "
                f"({R} rays per GPU), NCCL gradient allreduce")
    else:
        what = f"configs[1]: 800x800 Blender-shape scene, {R}-ray batch per GPU"
    mode = {"fp32": "fp32 training (CUDA cores)", "fp32_tc": "fp32 training (fp32-accurate bf16x3 tensor-core MLP)",
            "bf16": "bf16 tensor-core MLP"}[precision]
    return {"workload": f"{what}, {N_SAMPLES}+{N_SAMPLES} samples, 8x256 MipNeRF MLP ({n_params} params), {mode}",
            "rays_per_gpu": R, "global_batch": world * R, "precision": precision,
            "parallelism": f"ray-sharded dp{world}" if world > 1 else "single GPU",
            "l2": "per-step working set (activation cache, >4 GB) exceeds the 126 MB L2; no flush needed"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        out = self.window(t0, t1)
        self.close()
        return out

    def close(self):
        if self.proc:
            self.proc.terminate()
            self.proc = None

    def window(self, t0, t1):
        """clocks / board power / throttle reasons of the samples taken in [t0, t1] (the sampler keeps running)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for line in rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    continue
                try:
                    pw.append(float(f[2]))
                except Exception:
                    pass
                for n, v in zip(names, f[3:7]):
                    if v.lower() == "active":
                        reasons.add(n)
            return sm, mx, pw, reasons

        sm, mx, pw, reasons = collect([line for ts, line in self.rows if t0 - 0.05 <= ts <= t1 + 0.05])
        if not sm:  # region shorter than the sampling period: take every sample we have
            sm, mx, pw, reasons = collect([line for _, line in self.rows])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm


def cpu_port_rate(target_seconds, rank0=True):
    """rays/s of the CPU port (oracle) for one training step (gradient + Adam) on a bounded ray sample."""
    from oracle import oracle as orc  # the ONLY use of oracle/ in this file: the CPU baseline being measured
    from nerf_or_nothing_b200.scene import synthetic_rays

    ocfg = orc.default_config(**model_kw())
    params = orc.init_params(ocfg, 7)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would turn the
    # multi-threaded CPU port into a single-threaded one (only rank 0 runs this arm, the other ranks exit)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if orc.max_threads() < avail:
        orc.set_threads(avail)
    threads = orc.max_threads()

    def run(n, seed):
        rays, pix = synthetic_rays(n, width=800, height=800, seed=seed)
        u = np.stack([orc.sampling_uniforms(99, 0, lv, 0, n, N_SAMPLES + 1) for lv in range(2)])
        t0 = time.perf_counter()
        o = orc.train_gradient(ocfg, params, rays, pix, u, prec="f32")
        orc.adam_step(params, o["grads"], np.zeros_like(params), np.zeros_like(params), 5e-4, 1, 0, prec="f32")
        return time.perf_counter() - t0

    probe_n = max(threads, 8)
    dt = run(probe_n, 1)
    n = int(max(probe_n, min(4096, target_seconds * probe_n / max(dt, 1e-3))))
    n = max(threads, n // threads * threads)
    return run, n, threads


def reference_arm(args):
    """`--impl reference`: the reference's CPU path (C# restated in C, oracle/ — kind "port") on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 150.0 / max(1, args.steps + args.warmup)
    run, n, threads = cpu_port_rate(min(12.0, budget))
    for i in range(args.warmup):
        run(n, 10 + i)
    t = [run(n, 100 + i) for i in range(args.steps)]
    ms = 1e3 * float(np.mean(t))
    value = n / (ms / 1e3)
    sample = f"{n} of {RAYS_PER_GPU} rays per step (same config), gradient + Adam, fp32"
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(config_dict(args.rays, max(1, args.gpus), args.precision), rays_timed_per_step=n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ roofline arithmetic (pure)

GEMM_FAMILIES = ("mlp_fwd_gemm", "mlp_dgrad_gemm", "mlp_wgrad_gemm")


def kernel_table(prof, steps, R, S, precision, hbm_peak, tc_peak):
    """Per-kernel-family figures from the in-stream CUDA-event profile {name: (total ms, launches)}.

    SURVEY §8(d): the MLP GEMM families are bounded by the TENSOR pipe and are reported against it — `achieved` =
    algorithmic FLOPs (2 M sum(out x in); bias / activation / encoding FLOPs not counted) / time, `frac` = achieved /
    measured dense bf16 peak.  Side fields say what else loads the kernel: `hbm_gbs` / `frac_hbm` (the activation /
    gradient planes it must stream, `gemm_bytes`) and, in the fp32-accurate mode, `frac_tensor_issued` (every product is
    three bf16 MMAs).  Everything else is bounded by HBM: algorithmic bytes / time against the measured copy bandwidth."""
    work, _ = algorithmic_work(R, S)
    gbytes = gemm_bytes(R, S, precision)
    out = {}
    for name, (ms, nl) in prof.items():
        per_step_ms = ms / steps
        k = {"ms_per_step": round(per_step_ms, 4), "launches_per_step": nl / steps, "achieved": None, "unit": "GB/s", "bound": "hbm",
             "frac": None}
        if per_step_ms > 0:
            if name in GEMM_FAMILIES:
                tf = work[name][1] / 1e12 / (per_step_ms / 1e3)
                gbs = gbytes[name] / 1e9 / (per_step_ms / 1e3)
                k.update({"achieved": round(tf, 2), "unit": "TFLOP/s", "bound": "tensor", "frac": round(tf / tc_peak, 4),
                          "hbm_gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)})
                if precision == "fp32_tc" or (precision == W16_MODE and name != "mlp_wgrad_gemm"):
                    k["frac_tensor_issued"] = round(3 * tf / tc_peak, 4)
            else:
                nbytes = gbytes.get(name) or work.get(name, ("GB/s", 0))[1]
                if nbytes:
                    gbs = nbytes / 1e9 / (per_step_ms / 1e3)
                    k.update({"achieved": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)})
        out[name] = k
    return out


def pick_roofline(kernels, ms_step_prof, traffic_table, hbm_peak, tc_peak, peak_src):
    """`roofline` of the JSON line: the kernel family with the largest share of the step, against ITS roof (§8(d))."""
    top = max((k for k in kernels if kernels[k]["achieved"] is not None), key=lambda k: kernels[k]["ms_per_step"])
    tk = kernels[top]
    r = {"kernel": top, "bound": tk["bound"], "achieved": tk["achieved"], "peak": tc_peak if tk["bound"] == "tensor" else hbm_peak,
         "unit": tk["unit"], "frac": tk["frac"], "traffic": (traffic_table or {}).get(top, {}).get("dram_bytes_per_launch"),
         "peak_source": peak_src, "avg_launch_ms": round(tk["ms_per_step"] / max(1.0, tk["launches_per_step"]), 5),
         "share_of_step": round(tk["ms_per_step"] / ms_step_prof, 4)}
    for side in ("hbm_gbs", "frac_hbm", "frac_tensor_issued"):
        if side in tk:
            r[side] = tk[side]
    if "frac_tensor_issued" in tk:
        r["note"] = ("fp32-accurate mode: `achieved`/`frac` count ALGORITHMIC FLOPs; each product is three bf16 MMAs (hi*hi + lo*hi + "
                     "hi*lo), so the tensor pipe's load is frac_tensor_issued of the measured bf16 peak")
    return r


def traffic_table(precision):
    tpath = ROOT / "profiles" / "ncu_traffic.json"  # dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full`
    return json.loads(tpath.read_text()).get(precision, {}) if tpath.exists() else {}


# ------------------------------------------------------------------------------------------------ GPU arm


class Job:
    """Process-wide state of a GPU arm: rank / world, torch.distributed, ONE clock sampler for every timed window."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            args.gpus = self.world
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.clocks = ClockSampler(self.local)
        self.clocks.start()
        time.sleep(0.25)

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        self.clocks.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def make_batches(job, R, pool=4):
    from nerf_or_nothing_b200.scene import synthetic_rays

    host_batches, dev_batches = [], []
    for b in range(pool):
        rays, pix = synthetic_rays(R, width=800, height=800, seed=2024 + 1000 * job.rank + b)
        hb = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"], pix)
        host_batches.append(hb)
        dev_batches.append(tuple(job.torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in hb))
    return host_batches, dev_batches


def measure_train(job, args, precision, R, host_batches, dev_batches, want_e2e=True, want_dataset=True, engine_flags=None):
    """configs[1] / configs[2] step at R rays per GPU in `precision`.  Region 1: the headline — K steps, device-resident
    inputs, CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.  Region 2: the same K
    steps with the per-kernel in-stream events.  Then e2e (host arrays in, loss out) and the resident-dataset loop."""
    import nerf_or_nothing_b200 as nb

    torch = job.torch
    S, pool = N_SAMPLES, len(dev_batches)
    cfg = nb.default_config(n_rays=R, precision=nb.PRECISIONS[precision], device=job.local,
                            engine_flags=args.engine_flags if engine_flags is None else engine_flags, **model_kw())
    model = nb.AcceleratedMipNeRF(cfg)
    opt = nb.AcceleratedAdamOptimizer(model.GetLayerSizes(), device=job.local)
    if job.world > 1:
        from nerf_or_nothing_b200 import dist as nd

        nd.attach(model, device="cuda")  # rank 0's NCCL unique id -> every rank; the library allreduces [gradient | sum(lm) | losses]
    stream = torch.cuda.ExternalStream(model.stream())
    lr = nb.learning_rate_decay(1000)

    def dev_step(i):
        model.train_step_dev(opt, *dev_batches[i % pool], R, lr, want_loss=False)

    for i in range(args.warmup):
        dev_step(i)

    def timed_region(profile):
        model.set_profiling(profile)
        job.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        t0 = time.time()
        e0.record(stream)
        for i in range(args.steps):
            dev_step(i)
        e1.record(stream)
        model.synchronize()
        t1 = time.time()
        job.sync_all()
        return job.max_over_ranks(e0.elapsed_time(e1)) / args.steps, model.launch_count() - l0, t0, t1

    ms_step, launches, t0, t1 = timed_region(False)
    ms_step_prof, _, _, t1 = timed_region(True)
    clk = job.clocks.window(t0, t1)
    prof = model.read_profile()
    model.set_profiling(False)
    res = {"ms_step": ms_step, "ms_step_prof": ms_step_prof, "launches": int(launches), "clocks": clk, "prof": prof,
           "value": job.world * R / (ms_step / 1e3), "n_levels": cfg.n_levels}

    if want_e2e:  # host arrays in, loss out, every step
        for i in range(min(3, args.warmup)):
            model.train_step(opt, *host_batches[i % pool], lr, want_loss=True)
        job.sync_all()
        w0 = time.perf_counter()
        loss = 0.0
        for i in range(args.steps):
            loss = model.train_step(opt, *host_batches[i % pool], lr, want_loss=True)
        model.synchronize()
        ms = job.max_over_ranks((time.perf_counter() - w0) * 1e3)
        res.update({"e2e_value": job.world * R / (ms / args.steps / 1e3), "loss": loss})
    if want_dataset:  # the same loop fed from the device-resident BinDataset
        from nerf_or_nothing_b200.scene import pack_records

        keys = ("origins", "directions", "radii", "nears", "fars", "loss_mults")
        ds = nb.BinDataset(np.concatenate([pack_records(dict(zip(keys, hb[:6])), hb[6]) for hb in host_batches]), device=job.local)
        for i in range(min(3, args.warmup)):
            model.train_step_dataset(opt, ds, R, 2024, lr, want_loss=True)
        job.sync_all()
        w0 = time.perf_counter()
        for i in range(args.steps):
            model.train_step_dataset(opt, ds, R, 2024, lr, want_loss=True)
        model.synchronize()
        ms = job.max_over_ranks((time.perf_counter() - w0) * 1e3)
        res["resident_value"] = job.world * R / (ms / args.steps / 1e3)
        del ds
    if job.world > 1:
        # replicated training is only correct if every rank holds the same bits: compare a checksum of the parameters
        p = model.get_params()
        mine = torch.tensor([int(np.frombuffer(p.tobytes(), np.uint32).astype(np.uint64).sum() & 0x7FFFFFFFFFFFFFFF)], device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(job.world)]
        job.dist.all_gather(allv, mine)
        res["dp_check"] = bool(all(int(v.item()) == int(allv[0].item()) for v in allv))
        if not res["dp_check"]:
            raise RuntimeError("data-parallel check failed: parameters differ between ranks after the timed steps")
    model.close(); opt.close()
    torch.cuda.empty_cache()
    return res


def measure_render(job, args, precision, steps=None):
    """configs[3]: one 800x800 view = 640 000 rays, 128+128 samples, forward only, rays split over the ranks, no collective."""
    import nerf_or_nothing_b200 as nb
    from nerf_or_nothing_b200 import dist as nd
    from nerf_or_nothing_b200.scene import synthetic_rays

    torch = job.torch
    steps = steps or args.steps
    total, chunk = 640000, 16384
    lo, hi = nd.shard_range(total, job.rank, job.world)
    n = hi - lo
    cfg = nb.default_config(n_rays=chunk, precision=nb.PRECISIONS[precision], device=job.local, engine_flags=args.engine_flags, **model_kw())
    model = nb.AcceleratedMipNeRF(cfg)
    rays, _ = synthetic_rays(n, width=800, height=800, n_views=1, seed=7 + job.rank)
    hb = [rays[k] for k in ("origins", "directions", "radii", "nears", "fars")]
    db = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in hb]
    rgb, depth, acc = torch.empty(n, 3, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    stream = torch.cuda.ExternalStream(model.stream())
    for _ in range(3):  # >= 3 untimed passes
        model.render_dev(*db, n, rgb, depth, acc)
    model.set_profiling(True)
    job.sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = model.launch_count()
    t0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        model.render_dev(*db, n, rgb, depth, acc)
    e1.record(stream)
    model.synchronize()
    t1 = time.time()
    job.sync_all()
    prof = model.read_profile()
    launches = model.launch_count() - l0
    ms = job.max_over_ranks(e0.elapsed_time(e1)) / steps
    w0 = time.perf_counter()
    model.render(*hb)
    ms_e2e = job.max_over_ranks((time.perf_counter() - w0) * 1e3)
    # the same pixels rendered straight from the camera pose: rays generated on the device (Dataset.GenerateRays), image to the host
    from nerf_or_nothing_b200.scene import view_pose

    c2w, focal = view_pose(0, width=800)
    model.render_view(c2w, focal, 800, 800, first_pixel=lo, n_pixels=n)
    w0 = time.perf_counter()
    model.render_view(c2w, focal, 800, 800, first_pixel=lo, n_pixels=n)
    ms_pose = job.max_over_ranks((time.perf_counter() - w0) * 1e3)
    clk = job.clocks.window(t0, t1)
    hbm_peak, tc_peak, peak_src = peaks()
    work, _ = algorithmic_work(n, N_SAMPLES)
    fwd_ms = prof.get("mlp_fwd_gemm", (0, 0))[0] / steps
    ach = work["mlp_fwd_gemm"][1] / 1e12 / (fwd_ms / 1e3) if fwd_ms else None
    model.close()
    del db, rgb, depth, acc
    torch.cuda.empty_cache()
    arith = {"bf16": "bf16 products, fp32 accumulate", "fp32": "fp32 FFMA",
             "fp32_tc": ("fp32-accurate split on tcgen05: fp16 x fp16 plus two E4M3 correction products per term onto one fp32 accumulator"
                         if not (args.engine_flags & 256) else "fp32-accurate split on tcgen05: three bf16 products per term")}[precision]
    return {"value": total / (ms / 1e3), "unit": UNIT, "ms_per_image": ms, "steps": steps, "precision": precision, "arithmetic": arith, "rays": total,
            "rays_per_gpu": n, "chunk_rays": chunk, "gpu_launches": int(launches), "clocks": clk,
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * 9 * 4, "d2h_bytes_per_step": n * 5 * 4},
            "from_pose": {"value": total / (ms_pose / 1e3), "unit": UNIT, "h2d_bytes_per_step": 48, "d2h_bytes_per_step": n * 5 * 4,
                          "note": "nerf_mipnerf_render_view: the view's rays generated on the device from its 3x4 pose"},
            "roofline": {"kernel": "mlp_fwd_gemm", "bound": "tensor", "achieved": None if ach is None else round(ach, 2), "peak": tc_peak,
                         "unit": "TFLOP/s", "frac": None if ach is None else round(ach / tc_peak, 4), "traffic": None, "peak_source": peak_src,
                         "share_of_step": round(fwd_ms / ms, 4) if ms else None},
            "kernels": {k: {"ms_per_step": round(v[0] / steps, 4), "launches_per_step": v[1] / steps} for k, v in prof.items()}}


def measure_compositing(job, args, sizes=(65536, 262144), steps=None, warmup=None):
    """The two HBM-bound per-ray kernels alone, against the copy-bandwidth roofline (north_star: >= 70 % of HBM peak):
    `nerf_volumetric_rendering_async` / `nerf_volumetric_rendering_gradient_async` launched back to back on one stream,
    S = 128 samples per ray, R rays per launch, inputs larger than the 126 MB L2 (R*S*20 B >= 168 MB), timed with CUDA events
    on that stream.  Two forms: `activated` = the reference kernels' contract (.cu:318-344, 362-402: sigma and rgb in),
    `raw` = what the training step runs (softplus / sigmoid and their derivatives fused).  Algorithmic bytes: SURVEY §8(d),
    fwd 24 B/sample + 32 B/ray, bwd 36 B/sample + 24 B/ray."""
    import ctypes as C

    import nerf_or_nothing_b200 as nb

    torch = job.torch
    steps, warmup = steps or args.steps, warmup or args.warmup
    hbm_peak, _, peak_src = peaks()
    S = N_SAMPLES
    lib = nb.lib()
    st = torch.cuda.Stream()
    sp = C.c_void_p(st.cuda_stream)
    P = lambda x: C.c_void_p(x.data_ptr())
    cells = []
    launches = 0
    t_begin = time.time()
    for R in sizes:
        gen = torch.Generator(device="cuda").manual_seed(R)
        raw_rgb = torch.randn(R, S, 3, device="cuda", generator=gen) * 2.0
        raw_den = torch.randn(R, S, device="cuda", generator=gen) * 3.0 - 1.0
        act_rgb, act_den = torch.sigmoid(raw_rgb), torch.nn.functional.softplus(raw_den)
        t = torch.sort(torch.rand(R, S + 1, device="cuda", generator=gen) * 4.0 + 2.0, dim=1).values.contiguous()
        d = torch.randn(R, 3, device="cuda", generator=gen)
        g = torch.randn(R, 3, device="cuda", generator=gen)
        comp, depth, acc, w = (torch.empty(R, 3, device="cuda"), torch.empty(R, device="cuda"), torch.empty(R, device="cuda"),
                               torch.empty(R, S, device="cuda"))
        d_rgb, d_den = torch.empty(R, S, 3, device="cuda"), torch.empty(R, S, device="cuda")
        torch.cuda.synchronize()
        for form, rgb, den, raw in (("activated", act_rgb, act_den, 0), ("raw", raw_rgb, raw_den, 1)):
            def fwd():
                nb.check(lib.nerf_volumetric_rendering_async(P(rgb), P(den), P(t), P(d), P(comp), P(depth), P(acc), P(w), R, S, 1,
                                                             raw, 0.0, 0.0, sp))

            def bwd():
                nb.check(lib.nerf_volumetric_rendering_gradient_async(P(g), P(rgb), P(den), P(t), P(d), P(d_rgb), P(d_den), R, S, 1, 0,
                                                                      raw, 0.0, 0.0, sp))

            for name, fn, nbytes in (("composite_fwd", fwd, R * S * 24 + R * 32), ("composite_bwd", bwd, R * S * 36 + R * 24)):
                for _ in range(warmup):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st.synchronize()
                e0.record(st)
                for _ in range(steps):
                    fn()
                e1.record(st)
                st.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / steps
                launches += steps
                gbs = nbytes / 1e9 / (us / 1e6)
                cells.append({"kernel": name, "form": form, "rays": R, "samples": S, "us_per_launch": round(us, 2),
                              "algorithmic_bytes": nbytes, "achieved": round(gbs, 1), "unit": "GB/s", "frac": round(gbs / hbm_peak, 4)})
        assert torch.isfinite(comp).all() and torch.isfinite(d_den).all()
        del raw_rgb, raw_den, act_rgb, act_den, t, w, d_rgb, d_den
        torch.cuda.empty_cache()
    clk = job.clocks.window(t_begin, time.time())
    head = [c for c in cells if c["rays"] == max(sizes) and c["form"] == "raw"]
    worst = min(head, key=lambda c: c["frac"])
    return {"worst": worst, "cells": cells, "clocks": clk, "gpu_launches": launches, "steps": steps,
            "roofline": {"kernel": worst["kernel"], "bound": "hbm", "achieved": worst["achieved"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": worst["frac"], "traffic": None, "peak_source": peak_src}}


def train_record(res, args, R, precision, job):
    """roofline + kernel table of one measure_train() result."""
    hbm_peak, tc_peak, peak_src = peaks()
    kernels = kernel_table(res["prof"], args.steps, R, N_SAMPLES, precision, hbm_peak, tc_peak)
    roofline = pick_roofline(kernels, res["ms_step_prof"], traffic_table(precision), hbm_peak, tc_peak, peak_src)
    return kernels, roofline


def ours_arm(args):
    job = Job(args)
    R = args.rays if not args.global_batch else args.global_batch // job.world
    host_batches, dev_batches = make_batches(job, R)
    res = measure_train(job, args, args.precision, R, host_batches, dev_batches)
    kernels, roofline = train_record(res, args, R, args.precision, job)
    extra = {}
    if not args.no_extras:
        # the other configurations of BASELINE.json's metric, measured in the same process with their own clock windows:
        # configs[2] (bf16 tensor-core mode; at N ranks this is the N x R-ray global batch with the NCCL allreduce),
        # configs[3] (render of one 800x800 view split over the ranks) and, on one GPU, configs[4]'s compositing cells
        def mode_record(o, ok, orf):
            return {"value": o["value"], "unit": UNIT, "ms_per_step": o["ms_step"], "e2e": {"value": o["e2e_value"], "unit": UNIT},
                    "gpu_launches": o["launches"], "clocks": o["clocks"], "roofline": orf, "dp_check": o.get("dp_check"),
                    "kernels": {k: {f: v[f] for f in ("ms_per_step", "achieved", "unit", "frac")} for k, v in ok.items() if v["ms_per_step"] >= 0.02}}

        def guarded(name, fn):
            """a failing sub-record is reported as such and does not cost the headline line (every rank runs the same code path)"""
            try:
                return fn()
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"bench: sub-record {name} failed: {e}\n")
                return {"error": str(e)[:300]}

        def other_mode():
            other = "bf16" if args.precision != "bf16" else "fp32_tc"
            o = measure_train(job, args, other, R, host_batches, dev_batches, want_e2e=True, want_dataset=False)
            ok, orf = train_record(o, args, R, other, job)
            return other, mode_record(o, ok, orf)

        def w16_mode():
            # the headline mode with NERF_FLAG_WGRAD_FP16: wgrad operands as single fp16 planes (opt-in; accuracy in DESIGN.md section 4)
            w = measure_train(job, args, "fp32_tc", R, host_batches, dev_batches, want_e2e=True, want_dataset=False,
                              engine_flags=args.engine_flags | W16_FLAG)
            wk, wrf = train_record(w, args, R, W16_MODE, job)
            return dict(mode_record(w, wk, wrf), note="fp32-accurate forward (fp16 + E4M3 correction products) and dgrad chain (three-term bf16 products); "
                        "wgrad reads ONE fp16 plane per operand: not the default, its gradient meets 1e-4 on real steps only")

        extra["modes"] = {}
        r = guarded("modes", other_mode)
        if "error" in r:
            extra["modes"]["other"] = r
        else:
            extra["modes"][r[0]] = r[1]
        if args.precision == "fp32_tc" and not (args.engine_flags & W16_FLAG):
            extra["modes"][W16_MODE] = guarded(W16_MODE, w16_mode)
        extra["render"] = {p: guarded("render." + p, lambda p=p: {k: v for k, v in measure_render(job, args, p, steps=3).items() if k != "kernels"})
                           for p in ("bf16", "fp32_tc")}
        if job.world == 1:
            def comp():
                c = measure_compositing(job, args, sizes=(262144,), steps=20, warmup=5)
                return {"worst": c["worst"], "clocks": c["clocks"], "cells": c["cells"],
                        "note": "stand-alone launches, 262144 rays x 128 samples (inputs > L2); `worst` = slower of fwd/bwd in the raw form"}

            extra["compositing"] = guarded("compositing", comp)
    if job.rank != 0:
        job.close()
        return
    cpu_baseline = None
    if job.world == 1 and not args.no_cpu_baseline:
        run, n, threads = cpu_port_rate(15.0)
        dt = run(n, 7)
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n} of {R} rays, one training step (gradient + Adam), fp32, {dt:.1f} s"}
    out = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": job.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_step"], "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "fp32_tc": "f32 (bf16x3 tensor-core split)", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": config_dict(R, job.world, args.precision, args.global_batch),
        "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": R * 13 * 4, "d2h_bytes_per_step": (1 + res["n_levels"]) * 4},
        "resident_dataset": {"value": res["resident_value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": (1 + res["n_levels"]) * 4,
                             "note": "nerf_mipnerf_train_step_dataset: batch drawn and gathered on the device"},
        "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roofline, "kernels": kernels,
        "profile_region": {"ms_per_step": round(res["ms_step_prof"], 4),
                           "note": "per-kernel times come from a second region of the same K steps with in-stream CUDA events"},
        "cpu_baseline": cpu_baseline, "loss_last_step": res["loss"],
    }
    if job.world > 1:
        out["dp_check"] = res["dp_check"]  # parameters bit-identical on every rank after the timed steps
    out.update(extra)
    emit(out)
    job.close()


def render_arm(args):
    """configs[3] as a line of its own.  value = render rays/s with the rays resident on the device; e2e = host arrays in,
    host image out."""
    job = Job(args)
    r = measure_render(job, args, args.precision)
    if job.rank == 0:
        emit({"metric": "render rays/sec (forward only)", "value": r["value"], "unit": UNIT, "n_gpus": job.world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": r["ms_per_image"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": args.precision, "data": "synthetic",
              "config": {"workload": "configs[3]: full-image 800x800 render (640000 rays, 128+128 samples), forward only, rays split "
                                     f"across {job.world} GPU(s)", "rays_per_gpu": r["rays_per_gpu"], "chunk_rays": r["chunk_rays"],
                         "precision": args.precision, "l2": "per-chunk working set (>1 GB of encodings / heads) exceeds the 126 MB L2; no flush needed"},
              "e2e": r["e2e"], "from_pose": r["from_pose"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "roofline": r["roofline"],
              "kernels": r["kernels"]})
    job.close()


def sweep_arm(args):
    """configs[4]: MLP (depth x width) x batch sweep on one GPU — tensor-pipe throughput of the GEMM families and HBM GB/s of
    compositing fwd/bwd against the roofline.  One JSON line per cell (not the driver's headline line), each with the
    clocks sampled during its own timed region."""
    import nerf_or_nothing_b200 as nb
    from nerf_or_nothing_b200.scene import synthetic_rays

    job = Job(args)
    torch = job.torch
    hbm_peak, tc_peak, peak_src = peaks()
    cells = [(d, w, r) for (d, w) in ((4, 128), (8, 256), (8, 512)) for r in (4096, 16384, 65536)]
    for depth, width, R in cells:
        kw = dict(n_samples=N_SAMPLES, net_depth=depth, net_width=width, net_depth_condition=1, net_width_condition=width // 2,
                  skip_layer=4, deg_point=16, deg_view=4)
        try:
            cfg = nb.default_config(n_rays=R, precision=nb.PRECISIONS[args.precision], **kw)
            model = nb.AcceleratedMipNeRF(cfg)
            opt = nb.AcceleratedAdamOptimizer(model.GetLayerSizes())
            rays, pix = synthetic_rays(R, width=800, height=800, seed=1)
            hb = (rays["origins"], rays["directions"], rays["radii"], rays["nears"], rays["fars"], rays["loss_mults"], pix)
            db = tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in hb)
            for _ in range(3):
                model.train_step_dev(opt, *db, R, 1e-4)
            model.set_profiling(True)
            stream = torch.cuda.ExternalStream(model.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = max(3, min(20, int(0.4 / (2e-6 * R * (width / 256) ** 2 + 1e-3))))  # ~0.4 s per cell: enough nvidia-smi samples
            model.synchronize()
            t0 = time.time()
            e0.record(stream)
            for _ in range(steps):
                model.train_step_dev(opt, *db, R, 1e-4)
            e1.record(stream)
            model.synchronize()
            t1 = time.time()
            ms = e0.elapsed_time(e1) / steps
            prof = model.read_profile()
            P, Dd, Wc = 96, 27, width // 2
            lay = [(width, P, 0)] + [(width, width, P if (i % 4 == 0) else 0) for i in range(1, depth)] + [(Wc, width, Dd)]
            M = R * N_SAMPLES * 2
            fwd = 2 * M * sum(o * (a + b) for o, a, b in lay)
            dgr = 2 * M * sum(o * a for i, (o, a, b) in enumerate(lay) if i > 0)

            def tf(name, flops):
                t = prof.get(name, (0, 0))[0] / steps
                return round(flops / 1e12 / (t / 1e3), 1) if t > 0 else None

            def gb(name, nbytes):
                t = prof.get(name, (0, 0))[0] / steps
                return round(nbytes / 1e9 / (t / 1e3), 1) if t > 0 else None

            cf, cb = gb("composite_fwd", M * 24 + 2 * R * 32), gb("composite_bwd", M * 36 + 2 * R * 24)
            tfs = {k: tf(n, f) for k, n, f in (("fwd", "mlp_fwd_gemm", fwd), ("dgrad", "mlp_dgrad_gemm", dgr), ("wgrad", "mlp_wgrad_gemm", fwd))}
            emit(({"sweep": f"{depth}x{width}", "rays": R, "precision": args.precision, "ms_per_step": round(ms, 3), "steps": steps,
                   "train_rays_per_s": round(R / (ms / 1e3)), "chunk_launches_composite": prof.get("composite_fwd", (0, 0))[1] / steps,
                   "fwd_tflops": tfs["fwd"], "dgrad_tflops": tfs["dgrad"], "wgrad_tflops": tfs["wgrad"],
                   "tensor_frac": {k: None if v is None else round(v / tc_peak, 3) for k, v in tfs.items()},
                   "tensor_peak_tflops": tc_peak, "composite_fwd_gbs": cf, "composite_bwd_gbs": cb, "hbm_peak_gbs": hbm_peak,
                   "composite_fwd_frac": None if cf is None else round(cf / hbm_peak, 3),
                   "composite_bwd_frac": None if cb is None else round(cb / hbm_peak, 3), "peak_source": peak_src,
                   "clocks": job.clocks.window(t0, t1)}))
            model.close(); opt.close()
            del db
            torch.cuda.empty_cache()
        except Exception as e:  # a cell that does not fit is reported, not fatal
            emit({"sweep": f"{depth}x{width}", "rays": R, "error": str(e)[:200]})
    job.close()


def compositing_arm(args):
    job = Job(args)
    c = measure_compositing(job, args)
    worst = c["worst"]
    emit({"metric": "compositing fwd/bwd GB/s vs HBM roofline", "value": worst["achieved"], "unit": "GB/s", "n_gpus": 1,
          "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
          "config": {"workload": "configs[4] compositing cells: S=128, 65536 / 262144 rays per launch, inputs larger than L2 (no flush)",
                     "value_is": "the slower of fwd/bwd in the form the training step runs (raw), 262144 rays"},
          "roofline": c["roofline"], "gpu_launches": c["gpu_launches"], "clocks": c["clocks"], "cells": c["cells"]})
    job.close()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries the JSON result line(s) and nothing else: libraries that printf to fd 1 (NCCL prints its version banner
    there) are pointed at stderr for the rest of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NERF_BENCH_PRECISION", "fp32_tc"), choices=["fp32", "fp32_tc", "bf16"],
                    help="fp32_tc (default): fp32-accurate bf16x3 split on tcgen05; fp32: CUDA-core FFMA; bf16: configs[2] mode")
    ap.add_argument("--mode", default="train", choices=["train", "render", "sweep", "compositing"],
                    help="render: configs[3], forward only; sweep: configs[4]; compositing: the per-ray kernels alone vs the HBM roofline")
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="configs[2] strong scaling: total rays per step, split evenly over the ranks (overrides --rays)")
    ap.add_argument("--engine-flags", type=int, default=0, help="nerf_config.engine_flags (NERF_FLAG_*): A/B runs of alternative schedules")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline configuration (no modes / render / compositing sub-records)")
    ap.add_argument("--profiler-run", action="store_true", help="under ncu only: do not raise --warmup to 3 (the line printed is not a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" and not args.profiler_run else args.warmup
    if args.profiler_run:
        args.no_extras = True
    if args.impl == "reference":
        reference_arm(args)
    elif args.mode == "render":
        render_arm(args)
    elif args.mode == "sweep":
        sweep_arm(args)
    elif args.mode == "compositing":
        compositing_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
