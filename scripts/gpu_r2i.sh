#!/bin/bash
# round 2, call I: direct global stores of the activation planes vs shared-memory boxes + TMA stores (both precisions), A/B in one call
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "direct" -x > $out/${tag}_pytest_direct.log 2>&1; rc=$?; echo "pytest direct rc=$rc" | tee -a $out/${tag}_status.txt
tail -4 $out/${tag}_pytest_direct.log
if [ $rc -ne 0 ]; then echo "direct-store tests failed: stopping here"; exit 0; fi
for i in 1 2; do
for prec in fp32_tc bf16; do
timeout -s KILL 300 python bench.py --precision $prec --no-cpu-baseline --no-extras > $out/${tag}_${prec}_tma_$i.json 2> $out/${tag}_${prec}_tma_$i.err
timeout -s KILL 300 python bench.py --precision $prec --engine-flags 128 --no-cpu-baseline --no-extras > $out/${tag}_${prec}_direct_$i.json 2> $out/${tag}_${prec}_direct_$i.err
done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*_[12].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
true
