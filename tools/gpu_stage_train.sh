mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=14 run t60 600 python -m pytest tests/test_gpu_60_train.py -q -m gpu -x -s -k "attention_backward or train_grads or gradient"
TAILN=30 run tkb16 200 python tools/train_kernel_bench.py
TAILN=30 run tkb8 200 env TPAT_ATTN_BWD_WARPS=8 python tools/train_kernel_bench.py
TAILN=3 run bt_graph 200 python bench.py --mode train --steps 10 --warmup 3 --no-e2e
