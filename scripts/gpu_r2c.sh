#!/bin/bash
# round 2, call C: 2-CTA cluster weight multicast in the fp32-accurate fused kernels — targeted tests under a hard timeout, A/B bench.
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout -s KILL 200 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "multicast" -x -s > $out/${tag}_pytest_mc.log 2>&1; rc=$?; echo "pytest multicast rc=$rc" | tee -a $out/${tag}_status.txt
tail -15 $out/${tag}_pytest_mc.log
if [ $rc -ne 0 ]; then echo "multicast tests failed: stopping here"; exit 0; fi
timeout -s KILL 600 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_golden.py -q -m gpu -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -12
for i in 1 2; do
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-extras > $out/${tag}_bench_pair_$i.json 2> $out/${tag}_bench_pair_$i.err; echo "bench pair $i rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --engine-flags 64 --no-cpu-baseline --no-extras > $out/${tag}_bench_single_$i.json 2> $out/${tag}_bench_single_$i.err; echo "bench single $i rc=$?" | tee -a $out/${tag}_status.txt
done
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 > $out/${tag}_render_pair.json 2> $out/${tag}_render_pair.err
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags 64 > $out/${tag}_render_single.json 2> $out/${tag}_render_single.err
python - <<PY
import json
for f in ("bench_pair_1", "bench_single_1", "bench_pair_2", "bench_single_2", "render_pair", "render_single"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do echo "== $f"; tail -n 3 $f; done
