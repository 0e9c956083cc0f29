#!/bin/bash
# Staged GPU check: every stage in its own process (a CUDA fault poisons only that stage), logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "== $name rc=$? =="
  tail -n ${TAILN:-12} gpurun_out/$name.log
}
run t00 600 python -m pytest tests/test_gpu_00_kernels.py -q -m gpu -x
run t10a 600 python -m pytest tests/test_gpu_10_tensorcore.py -q -m gpu -k gemm
run t10b 600 python -m pytest tests/test_gpu_10_tensorcore.py -q -m gpu -k attention
run t20 1200 python -m pytest tests/test_gpu_20_forward.py -q -m gpu -s
run t25 900 python -m pytest tests/test_gpu_25_parity_protocol.py -q -m gpu -s
run t30 900 python -m pytest tests/test_gpu_30_ablation.py -q -m gpu -s
run t40 600 python -m pytest tests/test_gpu_40_frontend.py -q -m gpu -s
run t50 900 python -m pytest tests/test_gpu_50_multigpu.py -q -m gpu -s
run smoke 300 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 10 --warmup 3
