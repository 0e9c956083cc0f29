# MUFU offload experiment: every 4th / 8th exponential of attention_tc4 on the FMA pipe (variants built with -DA4_POLY_EVERY=4 | 8)
mkdir -p gpurun_out
for v in poly4 poly8; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_$v.so timeout 300 python -m pytest tests/test_gpu_13_attention_v4.py -x -q -m gpu -k "v4" > gpurun_out/poly_pytest_$v.log 2>&1; echo "$v pytest rc=$?"; tail -2 gpurun_out/poly_pytest_$v.log
done
for i in 1 2; do
  timeout 120 python tools/attn_bench.py 2>&1 | tail -1
  TPAT_LIB_PATH=$PWD/variants/libtpat_poly8.so timeout 120 python tools/attn_bench.py 2>&1 | tail -1
  TPAT_LIB_PATH=$PWD/variants/libtpat_poly4.so timeout 120 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/poly_attn_ab.txt
# fine-tune step with / without the new forward attention kernel (same box)
for v in "" "TPAT_ATTN_V4=0"; do
  env $v timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-e2e 2>/dev/null | tail -1 | cut -c1-200
done | tee gpurun_out/poly_train_ab.txt
