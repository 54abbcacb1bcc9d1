#!/bin/bash
# round 2, call H: double-buffered activation stores in the bf16 fused kernels — tests under a hard timeout, then bf16 lines
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "bf16" -x > $out/${tag}_pytest_bf16.log 2>&1; rc=$?; echo "pytest bf16 rc=$rc" | tee -a $out/${tag}_status.txt
tail -4 $out/${tag}_pytest_bf16.log
if [ $rc -ne 0 ]; then echo "bf16 tests failed: stopping here"; exit 0; fi
for i in 1 2; do
timeout -s KILL 300 python bench.py --precision bf16 --no-cpu-baseline --no-extras > $out/${tag}_bf16_$i.json 2> $out/${tag}_bf16_$i.err
done
timeout -s KILL 300 python bench.py --precision bf16 --global-batch 32768 --no-cpu-baseline --no-extras > $out/${tag}_config2_bf16_n1.json 2> $out/${tag}_config2.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
true
