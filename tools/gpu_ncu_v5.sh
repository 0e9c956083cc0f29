mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention_tc5_kernel|attention_tc4_kernel" --launch-skip 6 -c 2 -o gpurun_out/r02ad_attn_v4_v5 -f python tools/attn_ab_bench.py v4,v5 > gpurun_out/ncu_v5.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_v5.log
