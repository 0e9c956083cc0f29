mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_13_attention_v4.py -x -q -m gpu > gpurun_out/v5_pytest.log 2>&1; echo "v5 pytest rc=$?"; tail -5 gpurun_out/v5_pytest.log
timeout 300 python tools/attn_ab_bench.py old,v4,v5 > gpurun_out/v5_attn_ab.txt 2>&1; echo "ab rc=$?"; tail -8 gpurun_out/v5_attn_ab.txt
: > gpurun_out/v5_forward_ab.txt
for round in 1 2; do
  for v in "" "TPAT_ATTN_V5=1"; do
    env $v timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/v5_forward_ab.txt 2>> gpurun_out/v5_forward_ab.err
  done
done
cat gpurun_out/v5_forward_ab.txt
