# A/B against variants/libtpat_base.so = the library built from the PREVIOUS commit (git archive HEAD token-pruning-audio-transformer_b200/csrc include | tar -x -C /tmp/base; tools/build_variant.sh variants/libtpat_base.so from there)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_10_tensorcore.py tests/test_gpu_20_forward.py tests/test_gpu_25_parity_protocol.py -x -q -m gpu > gpurun_out/xp_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/xp_pytest.log
for i in 1 2; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_base.so timeout 120 python tools/attn_bench.py 2>&1 | tail -1
  timeout 120 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/xp_attn_ab.txt
: > gpurun_out/xp_forward_ab.txt
for round in 1 2; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_base.so timeout 200 python tools/forward_ab.py "base (r02ab)" >> gpurun_out/xp_forward_ab.txt 2>> gpurun_out/xp_forward_ab.err
  timeout 200 python tools/forward_ab.py "transpose column sums" >> gpurun_out/xp_forward_ab.txt 2>> gpurun_out/xp_forward_ab.err
done
cat gpurun_out/xp_forward_ab.txt
timeout 900 python tools/sweep.py --graph > gpurun_out/sweeps.md 2> gpurun_out/sweeps.err; echo "sweep rc=$?"; tail -40 gpurun_out/sweeps.md
