#!/bin/bash
# the driver's own N>1 command line (default flags: sub-records included) on 2 GPUs, the reference arm under torchrun, and the newest GPU tests
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
timeout -s KILL 300 python -m pytest tests/test_dataset_gpu.py tests/test_dp_gpu.py -q -m gpu -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
grep -E "passed|failed|error|FAILED" $out/${tag}_pytest.log | tail -5
( time timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_n2_default.json 2> $out/${tag}_bench_n2_default.err ) 2> $out/${tag}_time.txt; echo "bench N=2 default rc=$?" | tee -a $out/${tag}_status.txt
( time timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/${tag}_bench_n2_reference.json 2> $out/${tag}_bench_n2_reference.err ) 2>> $out/${tag}_time.txt; echo "reference arm N=2 rc=$?" | tee -a $out/${tag}_status.txt
grep real $out/${tag}_time.txt
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench_n2_default.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"]), "rays/s e2e", round(d["e2e"]["value"]), "dp_check", d.get("dp_check"), d["clocks"])
print("bf16", round(d["modes"]["bf16"]["ms_per_step"], 3), round(d["modes"]["bf16"]["value"]), d["modes"]["bf16"]["dp_check"], {p: (round(r["ms_per_image"], 1), round(r["value"])) for p, r in d["render"].items()}, "compositing" in d)
lines = [l for l in open("$out/${tag}_bench_n2_reference.json").read().splitlines() if l.strip()]
print("reference lines", len(lines), json.loads(lines[-1])["value"])
PY
tail -n 3 $out/${tag}_bench_n2_default.err
true
