#!/bin/bash
# One gpurun call: parity tests, bench lines and ncu captures of the current build.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- bash scripts/gpu_round.sh [tag]
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
# 1. the compositing stage tests first (new kernels), then the whole GPU suite
timeout 300 python -m pytest tests/test_stages_gpu.py -x -q -m gpu -k "volumetric" > $out/${tag}_pytest_stage.log 2>&1; echo "stage tests rc=$?" | tee -a $out/${tag}_status.txt
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -15 $out/${tag}_pytest.log
# 2. bench lines (no profiler)
timeout 300 python bench.py --mode compositing --steps 20 --warmup 5 > $out/${tag}_compositing.json 2> $out/${tag}_compositing.err; echo "bench compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout 400 python bench.py > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $out/${tag}_status.txt
cat $out/${tag}_compositing.json | cut -c1-1500
cut -c1-600 $out/${tag}_bench_fp32_tc.json
cut -c1-400 $out/${tag}_bench_bf16.json
# 3. ncu: full capture of the compositing kernels at 262144 rays (launches 9-16 of a 1+1-launch compositing run), then the
#    launch list of the default bench, then a full capture of one step's GEMM-family kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_composite --launch-skip 8 --launch-count 8 \
  -o $out/${tag}_ncu_compositing python bench.py --mode compositing --steps 1 --warmup 1 --profiler-run > $out/${tag}_ncu_compositing.log 2>&1; echo "ncu compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_fp32_tc.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $out/${tag}_status.txt
timeout 500 ncu --set full --clock-control none -k 'regex:k_tc_wgrad|k_mlp_fused|k_tc_gemm_persist' --launch-skip 120 --launch-count 40 \
  -o $out/${tag}_ncu_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a $out/${tag}_status.txt
ls -la $out | tail -20
