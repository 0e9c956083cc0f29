mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t60g 900 python -m pytest tests/test_gpu_60_train.py -q -m gpu -s -k "graphed or adamw or training_loop"
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
run btrain_ng 900 python bench.py --mode train --steps 10 --warmup 3 --no-graph
run btrain2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 10 --warmup 3
run t50 900 python -m pytest tests/test_gpu_50_multigpu.py -q -m gpu -s
