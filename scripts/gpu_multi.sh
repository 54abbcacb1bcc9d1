#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the data-parallel hardware tests, weak scaling in the bench's fp32-accurate mode, and
# configs[2] — bf16, 32768-ray GLOBAL batch split over the ranks (strong scaling) — each a JSON line with clocks + dp_check.
#   gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_multi.sh r02m 8
tag=${1:-r02m}; maxn=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $out/${tag}_gpus.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_dp_gpu.py -q -m gpu -s > $out/${tag}_pytest_dp.log 2>&1; echo "pytest dp rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED|2-rank" $out/${tag}_pytest_dp.log | tail -12
port=29500
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  port=$((port + 1))
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port"; fi
  timeout -s KILL 400 $L bench.py --gpus $n --no-cpu-baseline --no-extras > $out/${tag}_weak_fp32_tc_n$n.json 2> $out/${tag}_weak_n$n.err; echo "weak fp32_tc N=$n rc=$?" | tee -a $out/${tag}_status.txt
  port=$((port + 1))
  if [ $n -ne 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port"; fi
  timeout -s KILL 400 $L bench.py --gpus $n --precision bf16 --global-batch 32768 --no-cpu-baseline --no-extras > $out/${tag}_config2_bf16_n$n.json 2> $out/${tag}_config2_n$n.err; echo "configs[2] bf16 N=$n rc=$?" | tee -a $out/${tag}_status.txt
  port=$((port + 1))
  if [ $n -ne 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port"; fi
  timeout -s KILL 400 $L bench.py --gpus $n --mode render --precision bf16 --steps 5 > $out/${tag}_render_bf16_n$n.json 2> $out/${tag}_render_n$n.err; echo "render bf16 N=$n rc=$?" | tee -a $out/${tag}_status.txt
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*_n[0-9].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "N", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", d.get("scaling"), "dp_check", d.get("dp_check"),
              (d.get("kernels", {}).get("allreduce") or {}).get("ms_per_step"), d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do [ -s $f ] && { echo "== $f"; tail -n 2 $f; }; done
