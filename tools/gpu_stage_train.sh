mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run probe0 300 python tools/probes/epilogue_probe.py
TPAT_GEMM_NO_L2_PREFETCH=1 run probe1 300 python tools/probes/epilogue_probe.py
run t60 900 python -m pytest tests/test_gpu_60_train.py -q -m gpu -k "gemm_train or training_step_at"
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
