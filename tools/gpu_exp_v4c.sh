# A/B against variants/libtpat_base.so = the library built from the PREVIOUS commit (git archive HEAD token-pruning-audio-transformer_b200/csrc include | tar -x -C /tmp/base; tools/build_variant.sh variants/libtpat_base.so from there)
# full GPU suite (v4 default, trimmed two-pass main pass) + smoke + v4 trace + two-pass A/B against the HEAD build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu -s -k "test_other_vit_sizes" > gpurun_out/v4c_pytest_sizes.log 2>&1; echo "sizes rc=$?"; grep "logits err" gpurun_out/v4c_pytest_sizes.log
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/v4c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v4c_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/v4c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/v4c_smoke.log
for i in 1 2; do
  TPAT_LIB_PATH=$PWD/variants/libtpat_base.so timeout 120 python tools/attn_bench.py 2>&1 | tail -1
  timeout 120 python tools/attn_bench.py 2>&1 | tail -1
done | tee gpurun_out/v4c_attn_two_pass_ab.txt
timeout 300 python tools/probes/attn_trace_v4.py 513 1 > gpurun_out/v4c_trace_513.txt 2>&1; echo "trace rc=$?"; tail -75 gpurun_out/v4c_trace_513.txt
