# A/B against variants/libtpat_base.so = the library built from the PREVIOUS commit (git archive HEAD token-pruning-audio-transformer_b200/csrc include | tar -x -C /tmp/base; tools/build_variant.sh variants/libtpat_base.so from there)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_13_attention_v4.py -x -q -m gpu > gpurun_out/v4d_pytest.log 2>&1; echo "v4 pytest rc=$?"; tail -3 gpurun_out/v4d_pytest.log
for i in 1 2; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_base.so timeout 120 python tools/attn_bench.py 2>&1 | tail -1
  timeout 120 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/v4d_attn_ab.txt
: > gpurun_out/v4d_forward_ab.txt
for round in 1 2; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_base.so timeout 200 python tools/forward_ab.py "base (r02ab)" >> gpurun_out/v4d_forward_ab.txt 2>> gpurun_out/v4d_forward_ab.err
  timeout 200 python tools/forward_ab.py "pipelined v4" >> gpurun_out/v4d_forward_ab.txt 2>> gpurun_out/v4d_forward_ab.err
done
cat gpurun_out/v4d_forward_ab.txt
timeout 300 python tools/probes/attn_trace_v4.py 513 0 > gpurun_out/v4d_trace_513.txt 2>&1; echo "trace rc=$?"; tail -60 gpurun_out/v4d_trace_513.txt | head -45
