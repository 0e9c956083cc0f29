mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n 3 gpurun_out/$name.log | cut -c1-700; }
run t50_2gpu 600 python -m pytest tests/test_gpu_50_multigpu.py -q -m gpu -s
run binfer2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-eager-baseline
run bref2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --impl reference --steps 3 --warmup 1
run btrain2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --mode train --steps 10 --warmup 3
