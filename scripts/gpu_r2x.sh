#!/bin/bash
# NERF_FLAG_FP8_CORRECTIONS: tests, then render and training A/B in the same call
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "fp8_corrections" -s --durations=5 -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED|fp8-corr|Error|assert" $out/${tag}_pytest.log | tail -40
for fl in 256 0; do
  timeout -s KILL 200 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags $fl > $out/${tag}_render_$fl.json 2> $out/${tag}_render_$fl.err; echo "render flags=$fl rc=$?" | tee -a $out/${tag}_status.txt
done
for fl in 384 128 0; do
  timeout -s KILL 200 python bench.py --no-extras --no-cpu-baseline --engine-flags $fl > $out/${tag}_train_$fl.json 2> $out/${tag}_train_$fl.err; echo "train flags=$fl rc=$?" | tee -a $out/${tag}_status.txt
done
python - <<PY
import json
for n in ("render_256", "render_0", "train_384", "train_128", "train_0"):
    try:
        d = json.loads(open("$out/${tag}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["roofline"].get("frac"), d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
true
