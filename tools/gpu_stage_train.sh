mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t60a 240 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "attention_backward"
run t60s 1500 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "training_step or training_loop"
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
run binfer 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
