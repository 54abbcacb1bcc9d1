for fl in 0 64; do
  timeout -s KILL 200 python bench.py --mode render --precision fp32_tc --steps 3 --engine-flags $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('render flags=$fl', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done
for fl in 0 64; do
  timeout -s KILL 200 python bench.py --no-extras --no-cpu-baseline --engine-flags $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('train flags=$fl', round(d['ms_per_step'],3), {a:k[a]['ms_per_step'] for a in ('mlp_fwd_gemm','mlp_dgrad_gemm')}, d['clocks']['sm_mhz'])"
done
