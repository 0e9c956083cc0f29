mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_10_tensorcore.py -x -q -m gpu > gpurun_out/tmac_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/tmac_pytest.log
for i in 1 2; do
  TPAT_GEMM_TMA_STORE=0 timeout 200 python tools/kernel_bench.py 2>&1 | grep -E "N= (513|178) (qkv|fc1 )" | tr '\n' '|'; echo " <- per-lane stores"
  timeout 200 python tools/kernel_bench.py 2>&1 | grep -E "N= (513|178) (qkv|fc1 )" | tr '\n' '|'; echo " <- TMA stores"
done | tee gpurun_out/tmac_kernel_ab.txt
: > gpurun_out/tmac_forward_ab.txt
for round in 1 2 3; do
  for v in "TPAT_GEMM_TMA_STORE=0" ""; do
    env $v timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/tmac_forward_ab.txt 2>> gpurun_out/tmac_forward_ab.err
  done
done
cat gpurun_out/tmac_forward_ab.txt
