mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t60k 600 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "attention_backward or training_step_at"
TAILN=60 run trace 600 python tools/probes/attn_bwd_trace.py 513
run kb 600 python tools/train_kernel_bench.py
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
