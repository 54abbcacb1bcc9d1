#!/bin/bash
# NERF_FLAG_WGRAD_FP16: its tests, then an A/B of the bench step with and without the flag in the same call
tag=${1:-r02w}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $out/${tag}_gpu.txt
timeout -s KILL 600 python -m pytest tests/test_tc_gpu.py tests/test_bench_config_parity_gpu.py -q -m gpu -k "wgrad_fp16 and not loss_curve" -s --durations=5 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED|fp16-wgrad|whole step" $out/${tag}_pytest.log | tail -30
for i in 1 2; do
  timeout -s KILL 200 python bench.py --no-extras --no-cpu-baseline --engine-flags 128 > $out/${tag}_bench_w16_$i.json 2> $out/${tag}_bench_w16_$i.err; echo "bench w16 $i rc=$?" | tee -a $out/${tag}_status.txt
  timeout -s KILL 200 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench_default_$i.json 2> $out/${tag}_bench_default_$i.err; echo "bench default $i rc=$?" | tee -a $out/${tag}_status.txt
done
python - <<PY
import json
for n in ("w16_1", "default_1", "w16_2", "default_2"):
    try:
        d = json.loads(open("$out/${tag}_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
true
