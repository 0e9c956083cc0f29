mkdir -p gpurun_out
N=${1:-8}
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n 3 gpurun_out/$name.log | cut -c1-600; }
run binfer$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5
run btrain$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --mode train --steps 20 --warmup 3
run t50 600 python -m pytest tests/test_gpu_50_multigpu.py -q -m gpu -s
