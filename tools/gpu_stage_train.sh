mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t60w 240 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "wgrad"
run t60k 600 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "row_bwd or colsum or pool_norm or gemm_train or adamw"
run t60s 1500 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "training_step or training_loop"
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
run t50 900 python -m pytest tests/test_gpu_50_multigpu.py -q -m gpu -s
run btrain2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 10 --warmup 3
run binfer2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3
