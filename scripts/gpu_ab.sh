#!/bin/bash
# The B build is NOT kept in the tree: make it first, e.g.
#   NERF_NVCC_DEFS=-DNERF_LATE_SHIP=0 python -m nerf_or_nothing_b200.build --force && mkdir -p scratch && cp nerf_or_nothing_b200/libnerfb200.so scratch/libnerfb200_early.so && python -m nerf_or_nothing_b200.build --force
# generic A/B of two builds in one call: nerf_or_nothing_b200/libnerfb200.so (A) vs scratch/libnerfb200_early.so (B); $2 = extra bench flags
tag=${1:-r02ab}
out=gpurun_out
mkdir -p $out
cp nerf_or_nothing_b200/libnerfb200.so /tmp/lib_a.so
for i in 1 2; do
  for v in a b; do
    if [ $v = b ]; then cp scratch/libnerfb200_early.so nerf_or_nothing_b200/libnerfb200.so; else cp /tmp/lib_a.so nerf_or_nothing_b200/libnerfb200.so; fi
    timeout -s KILL 200 python bench.py --no-cpu-baseline --no-extras ${2:-} > $out/${tag}_bench_${v}_$i.json 2> $out/${tag}_bench_${v}_$i.err; echo "bench $v $i rc=$?" | tee -a $out/${tag}_status.txt
  done
done
cp /tmp/lib_a.so nerf_or_nothing_b200/libnerfb200.so
python - <<PY
import json
for i in (1, 2):
    for v in ("a", "b"):
        try:
            d = json.loads(open("$out/${tag}_bench_%s_%d.json" % (v, i)).read().strip().splitlines()[-1])
            k = d["kernels"]
            print(v, i, round(d["ms_per_step"], 3), {a: k[a]["ms_per_step"] for a in ("mlp_fwd_gemm", "mlp_dgrad_gemm", "mlp_wgrad_gemm", "mlp_bwd_heads")}, d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
        except Exception as e:
            print(v, i, "failed", e)
PY
