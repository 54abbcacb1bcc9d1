#!/bin/bash
# A/B call: stage tests of the compositing kernels, compositing bench, fp32_tc bench with the fused dgrad chain opted in.
tag=${1:-r01e}
out=gpurun_out
mkdir -p $out
timeout -s KILL 300 python -m pytest tests/test_stages_gpu.py tests/test_model_gpu.py -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -4 $out/${tag}_pytest.log
timeout -s KILL 300 python bench.py --mode compositing --steps 20 --warmup 5 > $out/${tag}_compositing.json 2> $out/${tag}_compositing.err; echo "bench compositing rc=$?" | tee -a $out/${tag}_status.txt
timeout -s KILL 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_fp32_tc.json 2> $out/${tag}_bench_fp32_tc.err; echo "bench fp32_tc rc=$?" | tee -a $out/${tag}_status.txt
NERF_FUSED_DGRAD_SPLIT=1 timeout -s KILL 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_fp32_tc_fusedchain.json 2> $out/${tag}_bench_fp32_tc_fusedchain.err; echo "bench fp32_tc fused dgrad chain rc=$?" | tee -a $out/${tag}_status.txt
python - <<PY
import json
for f in ("bench_fp32_tc", "bench_fp32_tc_fusedchain"):
    try:
        d = json.loads(open("$out/${tag}_" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.05}, d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
try:
    d = json.loads(open("$out/${tag}_compositing.json").read().strip().splitlines()[-1])
    for c in d["cells"]:
        print(c["kernel"], c["form"], c["rays"], c["us_per_launch"], "us", c["achieved"], "GB/s", c["frac"])
    print(d["clocks"])
except Exception as e:
    print("compositing unreadable", e)
PY
