mkdir -p gpurun_out
timeout 900 python tools/sweep.py --graph > gpurun_out/sweeps.md 2> gpurun_out/sweeps.err; echo "sweep rc=$?"; tail -30 gpurun_out/sweeps.md
timeout 400 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; echo "kb rc=$?"; tail -30 gpurun_out/kernel_bench.txt
