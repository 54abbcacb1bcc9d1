#!/bin/bash
# One gpurun call: the whole GPU suite, every bench line (with clocks), the reference kernels timed beside the new ones, and the ncu
# captures of the final build.  Everything lands in gpurun_out/ (big .ncu-rep files are exported to CSV on the box and deleted).
#   gpurun --timeout 2400 -- bash scripts/gpu_round2.sh [tag]
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -q -m gpu --durations=8 -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -8
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_status.txt; tail -1 $out/${tag}_smoke.log
timeout -s KILL 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --engine-flags 128 --no-cpu-baseline --no-extras > $out/${tag}_bench_fp32_tc_wgrad_fp16.json 2> $out/${tag}_bench_w16.err; echo "bench fp32_tc + NERF_FLAG_WGRAD_FP16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags 256 > $out/${tag}_render_fp32_tc_bf16x3.json 2> $out/${tag}_render_bf16x3.err; echo "render fp32_tc bf16x3 kernels rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --precision bf16 --no-cpu-baseline --no-extras > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --precision bf16 --global-batch 32768 --no-cpu-baseline --no-extras > $out/${tag}_config2_bf16_n1.json 2> $out/${tag}_config2_n1.err; echo "configs[2] N=1 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision bf16 --steps 5 > $out/${tag}_render_bf16.json 2> $out/${tag}_render_bf16.err; echo "render bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode render --precision fp32_tc --steps 3 > $out/${tag}_render_fp32_tc.json 2> $out/${tag}_render_fp32_tc.err; echo "render fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --mode compositing --steps 20 --warmup 5 > $out/${tag}_compositing.json 2> $out/${tag}_compositing.err; echo "compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 400 python bench.py --mode sweep --precision bf16 > $out/${tag}_sweep_bf16.jsonl 2> $out/${tag}_sweep_bf16.err; echo "sweep bf16 rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 400 python bench.py --mode sweep --precision fp32_tc > $out/${tag}_sweep_fp32_tc.jsonl 2> $out/${tag}_sweep_fp32_tc.err; echo "sweep fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 400 python scripts/time_reference_kernels.py --out $out/${tag}_reference_kernel_times.json > $out/${tag}_reference_kernel_times.log 2>&1; echo "reference kernels rc=$?" | tee -a $out/${tag}_status.txt
tail -12 $out/${tag}_reference_kernel_times.log
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_fp32_tc_wgrad_fp16", "bench_bf16", "config2_bf16_n1", "render_bf16", "render_fp32_tc", "render_fp32_tc_bf16x3"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.02}, {k: d["roofline"].get(k) for k in ("kernel", "bound", "achieved", "frac", "traffic")}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
for f in ("sweep_bf16", "sweep_fp32_tc"):
    try:
        for line in open("$out/${tag}_" + f + ".jsonl"):
            d = json.loads(line)
            print(f, {k: d.get(k) for k in ("sweep", "rays", "ms_per_step", "train_rays_per_s", "tensor_frac", "composite_fwd_frac", "composite_bwd_frac", "error")}, (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "unreadable", e)
PY
# ncu: (1) launch list of the default bench command; (2) full capture of the GEMM-family kernels of one steady-state step
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_fp32_tc.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-run > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:k_tc_wgrad|k_mlp_fused|k_tc_gemm_persist' --launch-skip 48 --launch-count 14 \
  -o $out/${tag}_ncu_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-run > $out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a $out/${tag}_status.txt
[ -f $out/${tag}_ncu_gemm.ncu-rep ] && ncu -i $out/${tag}_ncu_gemm.ncu-rep --page raw --csv > $out/${tag}_ncu_gemm_raw.csv 2>/dev/null
rm -f $out/${tag}_ncu_gemm.ncu-rep
# (3) the render kernel of the fp32-accurate mode (fp16 + E4M3 correction products): two launches of a steady-state image
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mlp_fused_split' --launch-skip 60 --launch-count 2 \
  -o $out/${tag}_ncu_render python bench.py --mode render --precision fp32_tc --steps 1 > $out/${tag}_ncu_render.log 2>&1; echo "ncu render rc=$?" | tee -a $out/${tag}_status.txt
[ -f $out/${tag}_ncu_render.ncu-rep ] && ncu -i $out/${tag}_ncu_render.ncu-rep --page raw --csv > $out/${tag}_ncu_render_raw.csv 2>/dev/null
rm -f $out/${tag}_ncu_render.ncu-rep
for f in $out/${tag}_*.err; do [ -s $f ] && { echo "== $f"; tail -n 3 $f; }; done
du -sh $out
