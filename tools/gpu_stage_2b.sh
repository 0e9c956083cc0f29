mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_50_multigpu.py tests/test_gpu_20_forward.py -q -m gpu -k "multigpu or second_device or nccl or data_parallel or allreduce" > gpurun_out/t2gpu_skipped.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t2gpu_skipped.log
