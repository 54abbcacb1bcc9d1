#!/bin/bash
# The B build is NOT kept in the tree: make it first, e.g.
#   NERF_NVCC_DEFS=-DNERF_LATE_SHIP=0 python -m nerf_or_nothing_b200.build --force && mkdir -p scratch && cp nerf_or_nothing_b200/libnerfb200.so scratch/libnerfb200_early.so && python -m nerf_or_nothing_b200.build --force
# late ship (TMA stores of the second-half epilogue behind act_ready) vs the previous order: A/B in one call, then the tensor-core tests
tag=${1:-r02y2}
out=gpurun_out
mkdir -p $out
cp nerf_or_nothing_b200/libnerfb200.so /tmp/lib_late.so
for i in 1 2; do
  for v in late early; do
    if [ $v = early ]; then cp scratch/libnerfb200_early.so nerf_or_nothing_b200/libnerfb200.so; else cp /tmp/lib_late.so nerf_or_nothing_b200/libnerfb200.so; fi
    timeout -s KILL 200 python bench.py --no-cpu-baseline > $out/${tag}_bench_${v}_$i.json 2> $out/${tag}_bench_${v}_$i.err; echo "bench $v $i rc=$?" | tee -a $out/${tag}_status.txt
  done
done
cp /tmp/lib_late.so nerf_or_nothing_b200/libnerfb200.so
python - <<PY
import json
for i in (1, 2):
    for v in ("late", "early"):
        try:
            d = json.loads(open("$out/${tag}_bench_%s_%d.json" % (v, i)).read().strip().splitlines()[-1])
            k = d["kernels"]
            line = [v, i, round(d["ms_per_step"], 3), {a: k[a]["ms_per_step"] for a in ("mlp_fwd_gemm", "mlp_dgrad_gemm", "mlp_wgrad_gemm")}]
            for m, r in d["modes"].items():
                line += [m, round(r["ms_per_step"], 3), {a: r["kernels"][a]["ms_per_step"] for a in ("mlp_fwd_gemm", "mlp_dgrad_gemm")}]
            print(*line)
        except Exception as e:
            print(v, i, "failed", e)
PY
timeout -s KILL 900 python -m pytest tests/test_tc_gpu.py tests/test_golden.py tests/test_bench_config_parity_gpu.py -q -m gpu -k "not loss_curve" -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -3 $out/${tag}_pytest.log
