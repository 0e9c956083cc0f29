mkdir -p gpurun_out
timeout 600 python tools/train_kernel_bench.py > gpurun_out/train_kernel_bench.txt 2>&1; echo rc=$?; cat gpurun_out/train_kernel_bench.txt
