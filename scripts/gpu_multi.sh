#!/bin/bash
# Multi-GPU call (gpurun --gpus 8): the data-parallel hardware tests, weak scaling in the bench's fp32-accurate mode, configs[2] —
# bf16, 32768-ray GLOBAL batch split over the ranks (strong scaling) — and configs[3] on 8 GPUs; every line with clocks + dp_check.
#   gpurun --gpus 8 --timeout 1200 -- bash scripts/gpu_multi.sh r02m
tag=${1:-r02m}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $out/${tag}_gpus.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_dp_gpu.py -q -m gpu -s > $out/${tag}_pytest_dp.log 2>&1; echo "pytest dp rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED|2-rank" $out/${tag}_pytest_dp.log | tail -12
port=29500
run() {  # run <n> <outfile> <bench args...>
  n=$1; f=$2; shift 2; port=$((port + 1))
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port"; fi
  timeout -s KILL 300 $L bench.py --gpus $n --no-cpu-baseline --no-extras "$@" > $out/${tag}_$f.json 2> $out/${tag}_$f.err; echo "$f rc=$?" | tee -a $out/${tag}_status.txt
}
for n in 1 2 8; do run $n weak_fp32_tc_n$n; done
for n in 1 2 4 8; do run $n config2_bf16_n$n --precision bf16 --global-batch 32768; done
run 8 weak_bf16_n8 --precision bf16
run 8 render_bf16_n8 --mode render --precision bf16 --steps 5
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*_n[0-9].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "N", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", d.get("scaling"), "dp_check", d.get("dp_check"),
              "allreduce ms", (d.get("kernels", {}).get("allreduce") or {}).get("ms_per_step"), d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do if [ -s $f ]; then echo "== $f"; tail -n 2 $f; fi; done
true
