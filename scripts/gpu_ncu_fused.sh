#!/bin/bash
# ncu --set full of the fused training kernels (forward + dgrad chain) of one steady-state step, default flags
tag=${1:-r02q}
out=gpurun_out
mkdir -p $out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mlp_fused_split' --launch-skip 12 --launch-count 4 \
  -o $out/${tag}_ncu_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-run ${2:-} > $out/${tag}_ncu_fused.log 2>&1; echo "ncu rc=$?"
[ -f $out/${tag}_ncu_fused.ncu-rep ] && ncu -i $out/${tag}_ncu_fused.ncu-rep --page raw --csv > $out/${tag}_ncu_fused_raw.csv 2>/dev/null
rm -f $out/${tag}_ncu_fused.ncu-rep
python scripts/ncu_summary.py $out/${tag}_ncu_fused_raw.csv > $out/${tag}_ncu_fused_summary.json
python - <<PY
import json
for r in json.load(open("$out/${tag}_ncu_fused_summary.json")):
    print(r["kernel"][-40:], {k: round(v["value"], 2) for k, v in r.items() if isinstance(v, dict) and k in ("time", "tensor_pipe_active_pct", "issue_active_pct", "dram_pct_of_peak", "l2_throughput_pct", "sm_clock")})
PY
