#!/bin/bash
# round 2, call J: fp32-accurate fused training kernels — [32 x 64] store boxes (4-stage ring) vs [32 x 32] boxes, A/B in one call
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_tc_gpu.py tests/test_bench_config_parity_gpu.py -q -m gpu -k "fp32_tc" -x > $out/${tag}_pytest.log 2>&1; rc=$?; echo "pytest fp32_tc rc=$rc" | tee -a $out/${tag}_status.txt
tail -4 $out/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed: stopping here"; exit 0; fi
for i in 1 2 3; do
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-extras > $out/${tag}_big_$i.json 2> $out/${tag}_big_$i.err
timeout -s KILL 300 python bench.py --engine-flags 128 --no-cpu-baseline --no-extras > $out/${tag}_small_$i.json 2> $out/${tag}_small_$i.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*_[123].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
true
