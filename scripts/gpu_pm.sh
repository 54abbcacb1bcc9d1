#!/bin/bash
# NERF_FLAG_PAIR_MMA (512): bit-identity tests under a short timeout (a protocol bug would hang), then A/B
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
timeout -s KILL ${2:-120} python -m pytest tests/test_tc_gpu.py -q -m gpu -k "pair_mma" -x -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_status.txt
tail -15 $out/${tag}_pytest.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
if grep -q "passed" $out/${tag}_pytest.log && ! grep -q "failed" $out/${tag}_pytest.log; then
for fl in 512 0; do
  timeout -s KILL 200 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('render flags=$fl', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done
for fl in 512 0 640 128; do
  timeout -s KILL 200 python bench.py --no-extras --no-cpu-baseline --engine-flags $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('train flags=$fl', round(d['ms_per_step'],3), {a:k[a]['ms_per_step'] for a in ('mlp_fwd_gemm','mlp_dgrad_gemm')}, d['clocks']['sm_mhz'])"
done
fi
