"""CPU checks of bench.py: the algorithmic work model behind every roofline figure, the JSON contract
of the lines it prints (checked on the committed evidence of the final build and on a live run of the CPU reference
arm), and the arithmetic of the reported fractions.  No GPU and no CUDA call: importing bench.py must not need one."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402

PROFILES = ROOT / "profiles"


def _last_line(path):
    return json.loads(path.read_text().strip().splitlines()[-1])


def test_algorithmic_work_matches_survey_8d():
    R, S = 4096, 128
    work, n_params = bench.algorithmic_work(R, S)
    M = R * S * 2
    assert n_params == 546948                                        # the flat parameter buffer
    assert work["mlp_fwd_gemm"][1] + 2 * M * (256 * 1 + 128 * 3) == 1089536 * M   # 1 089 536 FLOP per sample forward, heads included
    assert work["mlp_wgrad_gemm"][1] == work["mlp_fwd_gemm"][1]
    # dgrad skips the gradient into the encodings: 105 216 FLOP per sample less than forward, minus the heads' share
    enc = 2 * (96 * 256 + 96 * 256 + 27 * 128)
    assert work["mlp_fwd_gemm"][1] - work["mlp_dgrad_gemm"][1] == enc * M
    assert work["composite_fwd"][1] == M * 24 + 2 * R * 32            # 25.4 MB per step at configs[1]
    assert work["composite_bwd"][1] == M * 36 + 2 * R * 24            # 37.9 MB
    assert round(work["composite_fwd"][1] / 1e6, 1) == 25.4 and round(work["composite_bwd"][1] / 1e6, 1) == 37.9
    assert work["adam"][1] == 546948 * 28 == 15314544                # 28 B per parameter
    assert all(u in ("GB/s", "TFLOP/s") for u, _ in work.values())


def test_gemm_byte_model_matches_design_md():
    R, S = 4096, 128
    f32, b16 = bench.gemm_bytes(R, S, "fp32_tc"), bench.gemm_bytes(R, S, "bf16")
    # forward 10.2 / 5.3 GB, dgrad 9.4 / 4.8 GB, wgrad 18.6 / 9.3 GB, heads backward 2.2 / 1.1 GB per step
    for name, a, b in (("mlp_fwd_gemm", 10.2, 5.3), ("mlp_dgrad_gemm", 9.4, 4.8), ("mlp_wgrad_gemm", 18.6, 9.3), ("mlp_bwd_heads", 2.2, 1.1)):
        assert abs(f32[name] / 1e9 - a) < 0.06, (name, f32[name])
        assert abs(b16[name] / 1e9 - b) < 0.06, (name, b16[name])
    # the fp32-accurate mode moves hi + lo planes: twice the plane bytes of bf16, the same bit planes
    assert f32["mlp_wgrad_gemm"] == 2 * b16["mlp_wgrad_gemm"]


def test_layer_table_is_the_reference_network():
    lay = bench.layer_table()  # 8 x 256 trunk with a skip at layer 4, density head, 128-wide view layer, rgb head
    assert len(lay) == 11
    assert lay[0] == (256, 96, 0) and lay[4] == (256, 256, 96) and lay[8] == (1, 256, 0) and lay[9] == (128, 256, 27) and lay[10] == (3, 128, 0)
    assert sum(o * (a + b) + o for o, a, b in lay) == 546948


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "clocks", "roofline")


@pytest.mark.parametrize("name", ["r01f_bench_fp32_tc.json", "r01f_bench_bf16.json", "r01g_n8_fp32_tc.json"])
def test_committed_bench_lines_keep_the_contract(name):
    d = _last_line(PROFILES / name)
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3 and "workload" in d["config"]
    assert "model" not in d["config"] and d["config"]["rays_per_gpu"] == 4096
    # value = the units all ranks processed / the timed region
    assert abs(d["value"] - d["n_gpus"] * 4096 / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "rays/s" and e["h2d_bytes_per_step"] == 4096 * 13 * 4 and e["d2h_bytes_per_step"] == 12
    assert 0 < e["value"] < d["value"]                              # host copies and the loss read-back cost something
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["traffic"] is not None
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["peak_source"].startswith(("measured", "fallback")) and r["peak"] > 0
    k = d["kernels"][r["kernel"]]
    assert abs(k["ms_per_step"] / d["profile_region"]["ms_per_step"] - r["share_of_step"]) < 1e-3
    if d["n_gpus"] == 1 and name == "r01f_bench_fp32_tc.json":
        b = d["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] >= 1 and b["unit"] == "rays/s" and b["value"] > 0 and "rays" in b["sample"]


def test_committed_launch_list_shares_agree_with_the_bench_line():
    """The ncu launch list of the same command (cold-cache, serialised) must give the top kernel families the same SHARE of the
    step as the in-stream CUDA events of the bench line."""
    import csv
    import re

    rows = list(csv.reader((PROFILES / "r01f_launches_fp32_tc.csv").open()))
    hdr = next(r for r in rows if "Kernel Name" in r)
    data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r is not hdr and r[0].isdigit()]
    names = [d["Kernel Name"] for d in data]
    adam = [i for i, n in enumerate(names) if "k_adam" in n]
    assert len(adam) >= 5
    step = data[adam[3] + 1:adam[4] + 1]
    assert len(step) == 54                                           # launches per step of the final build
    tot, fam = 0.0, {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for d in step:
        t = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[d["Metric Unit"]]
        tot += t
        n = re.sub(r"\(.*", "", d["Kernel Name"])
        if "k_mlp_fused_split<1>" in n: fam["fwd"] += t
        elif "k_mlp_fused_split<2>" in n: fam["dgrad"] += t
        elif "k_tc_wgrad" in n or "k_reduce_job" in n: fam["wgrad"] += t
    line = _last_line(PROFILES / "r01f_bench_fp32_tc.json")
    ms = line["profile_region"]["ms_per_step"]
    for f, key in (("fwd", "mlp_fwd_gemm"), ("dgrad", "mlp_dgrad_gemm"), ("wgrad", "mlp_wgrad_gemm")):
        assert abs(fam[f] / tot - line["kernels"][key]["ms_per_step"] / ms) < 0.02, (f, fam[f] / tot)


def test_compositing_line_meets_the_north_star_target():
    d = _last_line(PROFILES / "r01f_compositing.json")
    assert d["roofline"]["bound"] == "hbm" and not d["clocks"]["reasons"]
    cells = d["cells"]
    assert len(cells) == 8
    for c in cells:
        per_sample = 24 if c["kernel"] == "composite_fwd" else 36
        per_ray = 32 if c["kernel"] == "composite_fwd" else 24
        assert c["algorithmic_bytes"] == c["rays"] * (c["samples"] * per_sample + per_ray)
        assert abs(c["achieved"] - c["algorithmic_bytes"] / 1e9 / (c["us_per_launch"] / 1e6)) <= 1e-3 * c["achieved"]  # us rounded to 0.01
        assert c["rays"] * c["samples"] * 20 > 126e6                 # inputs larger than the L2: no flush needed
        assert c["frac"] >= 0.70, c                                  # north_star: >= 70 % of HBM peak on compositing fwd / bwd


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference's C# path) on a short run: one JSON line, same metric."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "rays/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    b = d["cpu_baseline"]
    assert b["kind"] == "port" and b["value"] == d["value"] and b["cores"] >= 1 and "rays" in b["sample"]
    assert d["config"]["workload"].startswith("configs[1]")
