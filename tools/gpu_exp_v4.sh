# experiment call: attention v4 tests + kernel A/B + whole-forward A/B (v4, L2 persistence)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_13_attention_v4.py -x -q -m gpu > gpurun_out/v4_pytest.log 2>&1; echo "v4 pytest rc=$?"; tail -4 gpurun_out/v4_pytest.log
timeout 300 python tools/attn_ab_bench.py old,v4 > gpurun_out/v4_attn_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/v4_attn_ab.txt | tail -8
: > gpurun_out/v4_forward_ab.txt
for round in 1 2; do
  for v in "" "TPAT_ATTN_V4=1" "TPAT_L2_PERSIST_MB=48" "TPAT_L2_PERSIST_MB=80" "TPAT_ATTN_V4=1 TPAT_L2_PERSIST_MB=64"; do
    env $v TPAT_DEBUG=1 timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/v4_forward_ab.txt 2>> gpurun_out/v4_forward_ab.err
  done
done
cat gpurun_out/v4_forward_ab.txt; grep -h "L2 persisting" gpurun_out/v4_forward_ab.err | sort | uniq -c
