# r02ab: evidence for the shipped build -- full GPU suite, smoke, benches, ncu launch list of one forward (+ DRAM bytes),
# ncu --set full of every kernel of interest (one launch each)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.log 2>&1; echo "b20 rc=$?"; tail -1 gpurun_out/bench_20.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-200
# launch list of one forward with DRAM bytes (eager launches so that every kernel is a separate ncu record)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_fwd_dram.csv \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
K='regex:attention_tc_kernel|attention_tc4_kernel|attn_delta8_kernel|attention_bwd_tc_kernel|dq_convert_sum_kernel|gemm_wgrad_tc_kernel|gemm_tc2_kernel|row_bwd_kernel|colsum8_kernel|partials_finish_kernel|layernorm_kernel|gather_layernorm_kernel|score_topk_kernel|patchify_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/pko_launches.csv python tools/profile_kernels_once.py > /dev/null 2>&1
n=$(grep -c "gpu__time_duration.sum" gpurun_out/pko_launches.csv); half=$((n / 2)); echo "launches=$n skip=$half"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip $half -o gpurun_out/r02ab_kernels -f python tools/profile_kernels_once.py > gpurun_out/pko_full.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/pko_full.log
ls -la gpurun_out/*.ncu-rep
