#!/bin/bash
# round 2, call G: 2-CTA weight multicast in the bf16 fused kernels — targeted tests under a hard timeout, then A/B lines.
tag=${1:-r02g}
out=gpurun_out
mkdir -p $out
timeout -s KILL 300 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "multicast" -x > $out/${tag}_pytest_mc.log 2>&1; rc=$?; echo "pytest multicast rc=$rc" | tee -a $out/${tag}_status.txt
tail -5 $out/${tag}_pytest_mc.log
if [ $rc -ne 0 ]; then echo "multicast tests failed: stopping here"; exit 0; fi
timeout -s KILL 900 python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -8
for i in 1 2; do
timeout -s KILL 300 python bench.py --precision bf16 --no-cpu-baseline --no-extras > $out/${tag}_bf16_pair_$i.json 2> $out/${tag}_bf16_pair_$i.err
timeout -s KILL 300 python bench.py --precision bf16 --engine-flags 64 --no-cpu-baseline --no-extras > $out/${tag}_bf16_single_$i.json 2> $out/${tag}_bf16_single_$i.err
timeout -s KILL 300 python bench.py --mode render --precision bf16 --steps 5 > $out/${tag}_render_bf16_pair_$i.json 2> $out/${tag}_render_pair_$i.err
timeout -s KILL 300 python bench.py --mode render --precision bf16 --steps 5 --engine-flags 64 > $out/${tag}_render_bf16_single_$i.json 2> $out/${tag}_render_single_$i.err
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
