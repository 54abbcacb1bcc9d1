#!/bin/bash
# round 2, call B: in-kernel encoder warps — targeted tests under a hard timeout first, then the whole suite and the bench lines.
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "inkernel or fused_forward_render" -x -s > $out/${tag}_pytest_enc.log 2>&1; echo "pytest enc rc=$?" | tee -a $out/${tag}_status.txt
tail -8 $out/${tag}_pytest_enc.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --durations=8 -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -12
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_status.txt; tail -2 $out/${tag}_smoke.log
timeout -s KILL 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --engine-flags 16 --no-cpu-baseline > $out/${tag}_bench_fp32_tc_encode_kernel.json 2> $out/${tag}_bench_ek.err; echo "bench encode-kernel rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_fp32_tc_encode_kernel"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"])
        m = d["modes"]["bf16"]
        print("   bf16", round(m["ms_per_step"], 3), round(m["value"]), {k: v["ms_per_step"] for k, v in m["kernels"].items()}, m["clocks"])
        for p, r in d["render"].items():
            print("   render", p, round(r["ms_per_image"], 2), "ms", round(r["value"]), "rays/s", r["roofline"]["frac"], r["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do echo "== $f"; tail -n 3 $f; done
