mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python tools/sweep.py --graph > gpurun_out/sweeps.md 2> gpurun_out/sweeps.err; echo "sweep rc=$?"; head -20 gpurun_out/sweeps.md
timeout 400 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; echo "kb rc=$?"; tail -30 gpurun_out/kernel_bench.txt
