#!/bin/bash
# round 2, call E: the fused kernels on the narrow network (4x128 / condition 64) — targeted tests under a hard timeout, then the sweep.
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "4x128" -x -s > $out/${tag}_pytest_narrow.log 2>&1; rc=$?; echo "pytest 4x128 rc=$rc" | tee -a $out/${tag}_status.txt
tail -25 $out/${tag}_pytest_narrow.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "narrow-net tests failed: stopping here"; exit 0; fi
timeout -s KILL 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_golden.py tests/test_bench_config_parity_gpu.py -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -12
timeout -s KILL 400 python bench.py --mode sweep --precision bf16 > $out/${tag}_sweep_bf16.jsonl 2> $out/${tag}_sweep_bf16.err; echo "sweep bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 400 python bench.py --mode sweep --precision fp32_tc > $out/${tag}_sweep_fp32_tc.jsonl 2> $out/${tag}_sweep_fp32_tc.err; echo "sweep fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("sweep_bf16", "sweep_fp32_tc"):
    try:
        for line in open("$out/${tag}_" + f + ".jsonl"):
            d = json.loads(line)
            print(f, {k: d.get(k) for k in ("sweep", "rays", "ms_per_step", "train_rays_per_s", "tensor_frac", "composite_fwd_frac", "composite_bwd_frac", "error")}, (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do [ -s $f ] && { echo "== $f"; tail -n 3 $f; }; done
