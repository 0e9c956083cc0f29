mkdir -p gpurun_out
for rep in 1 2 3; do
for mode in 0 proj all; do
  TPAT_GEMM_RES_REDUCE=$mode timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/bi_red_$mode.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/bi_red_$mode.log'):
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep mode $mode: bench', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
PY
done
done
