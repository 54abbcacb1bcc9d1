#!/bin/bash
# Short 8-GPU check of the shipped build: the driver's own N=8 command (default flags, sub-records included), N=1 on the same box,
# configs[2] (bf16, 32768-ray global batch) at N=8, and the fp16-wgrad option at N=8; every line with clocks + dp_check.
#   gpurun --gpus 8 --timeout 600 -- bash scripts/gpu_multi_final.sh r03b
tag=${1:-r03b}
out=gpurun_out
mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 200 python bench.py --gpus 1 --no-cpu-baseline --no-extras > $out/${tag}_weak_fp32_tc_n1.json 2> $out/${tag}_n1.err; echo "n1 rc=$?"
timeout -s KILL 300 $T --master-port 29511 bench.py --gpus 8 > $out/${tag}_bench_n8_default.json 2> $out/${tag}_n8.err; echo "n8 default rc=$?"
timeout -s KILL 200 $T --master-port 29512 bench.py --gpus 8 --precision bf16 --global-batch 32768 --no-cpu-baseline --no-extras > $out/${tag}_config2_bf16_n8.json 2> $out/${tag}_c2.err; echo "config2 n8 rc=$?"
timeout -s KILL 200 $T --master-port 29513 bench.py --gpus 8 --engine-flags 128 --no-cpu-baseline --no-extras > $out/${tag}_weak_fp32_tc_wgrad_fp16_n8.json 2> $out/${tag}_w16.err; echo "w16 n8 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        extra = ""
        if "modes" in d: extra = {k: (round(v["ms_per_step"], 3), round(v["value"]), v.get("dp_check")) for k, v in d["modes"].items()}, {p: round(r["ms_per_image"], 1) for p, r in d.get("render", {}).items()}
        print(f.split("/")[-1], "N", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s", d.get("scaling"), "dp_check", d.get("dp_check"), d["clocks"]["sm_mhz"], extra)
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $out/${tag}_*.err; do if [ -s $f ]; then echo "== $f"; tail -n 2 $f; fi; done
true
