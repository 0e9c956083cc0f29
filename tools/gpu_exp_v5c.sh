mkdir -p gpurun_out
timeout 300 python tools/probes/attn_trace_v4.py 513 1 > gpurun_out/v5c_trace_513.txt 2>&1; echo "trace rc=$?"; tail -150 gpurun_out/v5c_trace_513.txt
