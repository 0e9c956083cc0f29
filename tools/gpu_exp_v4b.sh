# full GPU suite with v4 as the default single-pass attention + L2 carve-out scan
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/v4b_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/v4b_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/v4b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/v4b_smoke.log
: > gpurun_out/v4b_forward_ab.txt
for round in 1 2; do
  for v in "" "TPAT_L2_PERSIST_MB=8" "TPAT_L2_PERSIST_MB=16" "TPAT_L2_PERSIST_MB=32" "TPAT_L2_PERSIST_MB=48" "TPAT_ATTN_V4=0"; do
    env $v timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/v4b_forward_ab.txt 2>> gpurun_out/v4b_forward_ab.err
  done
done
cat gpurun_out/v4b_forward_ab.txt
