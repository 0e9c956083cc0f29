mkdir -p gpurun_out
K='regex:attention_tc_kernel|attention_tc3_kernel|attn_delta8_kernel|attention_bwd_tc_kernel|dq_convert_sum_kernel|gemm_wgrad_tc_kernel|gemm_tc2_kernel|row_bwd_kernel|colsum8_kernel|partials_finish_kernel|layernorm_kernel|gather_layernorm_kernel|score_topk_kernel|patchify_kernel'
timeout 300 python tools/profile_kernels_once.py > gpurun_out/pko.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/pko.log
# count launches of one pass with a cheap ncu pass, then capture the second pass only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/pko_launches.csv python tools/profile_kernels_once.py > /dev/null 2>&1
n=$(grep -c "gpu__time_duration.sum" gpurun_out/pko_launches.csv); half=$((n / 2)); echo "launches=$n skip=$half"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip $half -o gpurun_out/r02z_kernels -f python tools/profile_kernels_once.py > gpurun_out/pko_full.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/pko_full.log
TPAT_ATTN_V3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3_kernel --launch-skip 1 -c 1 -o gpurun_out/r02z_attn_v3 -f python tools/attn_v3_bench.py > gpurun_out/v3_full.log 2>&1; echo "v3 rc=$?"
ls -la gpurun_out/*.ncu-rep
