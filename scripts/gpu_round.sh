#!/bin/bash
# One gpurun call: parity tests, bench lines and ncu captures of the current build.  Everything lands in gpurun_out/
# (kept under the 64 MiB that gpurun copies back: big .ncu-rep files are exported to CSV on the box and deleted).
#   gpurun --timeout 1200 -- bash scripts/gpu_round.sh [tag]
tag=${1:-r01f}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
# 1. the whole GPU suite (no -x: one failure must not hide the rest)
timeout 900 python -m pytest tests -q -m gpu --durations=5 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -12 $out/${tag}_pytest.log
# 2. bench lines (no profiler)
timeout 300 python bench.py --mode compositing --steps 20 --warmup 5 > $out/${tag}_compositing.json 2> $out/${tag}_compositing.err; echo "bench compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision bf16 --steps 3 > $out/${tag}_render_bf16.json 2> $out/${tag}_render_bf16.err; echo "render bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode sweep --precision bf16 > $out/${tag}_sweep_bf16.jsonl 2> $out/${tag}_sweep_bf16.err; echo "sweep bf16 rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_bf16", "render_bf16"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.05}, d["roofline"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
try:
    d = json.loads(open("$out/${tag}_compositing.json").read().strip().splitlines()[-1])
    for c in d["cells"]:
        print(c["kernel"], c["form"], c["rays"], c["us_per_launch"], "us", c["achieved"], "GB/s", c["frac"])
except Exception as e:
    print("compositing unreadable", e)
try:
    for line in open("$out/${tag}_sweep_bf16.jsonl"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("sweep", "rays", "ms_per_step", "train_rays_per_s", "fwd_tflops", "dgrad_tflops", "wgrad_tflops", "composite_fwd_gbs", "composite_bwd_gbs", "error")})
except Exception as e:
    print("sweep unreadable", e)
PY
# 3. ncu: full capture of the compositing kernels at 262144 rays (launches 9-16 of a 1+1-launch compositing run), the
#    launch list of the default bench, and a full capture of the first 14 GEMM-family kernels of a steady-state step (2 fused forward, 1 fused dgrad chain, 11 wgrad)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_composite --launch-skip 8 --launch-count 8 \
  -o $out/${tag}_ncu_compositing python bench.py --mode compositing --steps 1 --warmup 1 --profiler-run > $out/${tag}_ncu_compositing.log 2>&1; echo "ncu compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_fp32_tc.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $out/${tag}_status.txt
timeout 500 ncu --set full --clock-control none -k 'regex:k_tc_wgrad|k_mlp_fused|k_tc_gemm_persist' --launch-skip 52 --launch-count 14 \
  -o $out/${tag}_ncu_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a $out/${tag}_status.txt
for r in ncu_compositing ncu_gemm; do
  [ -f $out/${tag}_$r.ncu-rep ] && ncu -i $out/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
done
rm -f $out/${tag}_ncu_gemm.ncu-rep
du -sh $out; ls -la $out | tail -25
