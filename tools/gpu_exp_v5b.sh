mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_13_attention_v4.py -x -q -m gpu > gpurun_out/v5b_pytest.log 2>&1; echo "v5 pytest rc=$?"; tail -3 gpurun_out/v5b_pytest.log
for pct in 0 50 25 100; do
  echo "stagger pct $pct"
  TPAT_A5_STAGGER_PCT=$pct timeout 300 python tools/attn_ab_bench.py v4,v5 2>&1 | tail -6
done | tee gpurun_out/v5b_attn_ab.txt
: > gpurun_out/v5b_forward_ab.txt
for round in 1 2; do
  for v in "" "TPAT_ATTN_V5=1" "TPAT_ATTN_V5=1 TPAT_A5_STAGGER_PCT=25"; do
    env $v timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/v5b_forward_ab.txt 2>> gpurun_out/v5b_forward_ab.err
  done
done
cat gpurun_out/v5b_forward_ab.txt
