#!/bin/bash
# quick end-of-change check: the whole GPU suite, smoke, the default bench line and the reference arm
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --durations=5 -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -8
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_status.txt; tail -1 $out/${tag}_smoke.log
( time timeout -s KILL 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err ) 2> $out/${tag}_bench_time.txt; echo "bench rc=$?" | tee -a $out/${tag}_status.txt
( time timeout -s KILL 400 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err ) 2>> $out/${tag}_bench_time.txt; echo "reference arm rc=$?" | tee -a $out/${tag}_status.txt
cat $out/${tag}_bench_time.txt | grep real
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench_fp32_tc.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"])
print("bf16", round(d["modes"]["bf16"]["ms_per_step"], 3), round(d["modes"]["bf16"]["value"]), {p: (round(r["ms_per_image"], 1), round(r["value"])) for p, r in d["render"].items()}, d["compositing"]["worst"]["frac"])
r = json.loads(open("$out/${tag}_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", round(r["value"], 1), r["cpu_baseline"]["cores"], r["config"]["rays_timed_per_step"])
PY
true
