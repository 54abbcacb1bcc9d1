#!/bin/bash
# Short gpurun call: the GPU suite and the bench lines of the current build (no profiler).
#   gpurun --timeout 900 -- bash scripts/gpu_quick.sh [tag]
tag=${1:-r01d}
out=gpurun_out
mkdir -p $out
timeout -s KILL 600 python -m pytest tests -q -m gpu --durations=5 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -12 $out/${tag}_pytest.log
timeout -s KILL 300 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --precision bf16 --no-cpu-baseline > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision bf16 --steps 3 > $out/${tag}_render_bf16.json 2> $out/${tag}_render_bf16.err; echo "render bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 > $out/${tag}_render_fp32_tc.json 2> $out/${tag}_render_fp32_tc.err; echo "render fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_bf16", "render_bf16", "render_fp32_tc"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.05}, d["roofline"], d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $out/${tag}_*.err
