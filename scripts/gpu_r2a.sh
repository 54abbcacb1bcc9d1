#!/bin/bash
# round 2, call A: validate the quarter-granular schedule in isolation, then the whole GPU suite, the default bench line and the A/B.
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout -s KILL 240 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "quarters" -x > $out/${tag}_pytest_quarters.log 2>&1; echo "pytest quarters rc=$?" | tee -a $out/${tag}_status.txt
tail -5 $out/${tag}_pytest_quarters.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --durations=8 -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error" $out/${tag}_pytest.log | tail -5
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_status.txt; tail -2 $out/${tag}_smoke.log
timeout -s KILL 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --engine-flags 16 --no-cpu-baseline --no-extras > $out/${tag}_bench_fp32_tc_quarters.json 2> $out/${tag}_bench_fp32_tc_quarters.err; echo "bench quarters rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-extras > $out/${tag}_bench_fp32_tc_halves.json 2> $out/${tag}_bench_fp32_tc_halves.err; echo "bench halves rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags 16 > $out/${tag}_render_fp32_tc_quarters.json 2> $out/${tag}_render_q.err; echo "render q rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_fp32_tc_quarters", "bench_fp32_tc_halves"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["roofline"], d["clocks"])
        for k in ("modes", "render", "compositing"):
            if k in d: print("  ", k, json.dumps(d[k])[:1500])
    except Exception as e:
        print(f, "unreadable", e)
try:
    d = json.loads(open("$out/${tag}_render_fp32_tc_quarters.json").read().strip().splitlines()[-1])
    print("render fp32_tc quarters", d["ms_per_step"], d["value"], d["kernels"], d["clocks"])
except Exception as e:
    print("render unreadable", e)
PY
tail -3 $out/*.err | tail -30
