mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t60k 900 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -x -k "gemm_f32 or transpose or row_bwd or colsum or pool_norm or gemm_train or adamw"
run t60a 900 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "attention_backward"
run t60s 1500 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "training_step or training_loop"
run t10 600 python -m pytest tests/test_gpu_10_tensorcore.py -q -m gpu
