mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t60k 600 python -m pytest tests/test_gpu_60_train.py -q -m gpu -k "attention_backward or training_step_at"
run btrain 900 python bench.py --mode train --steps 10 --warmup 3
run fwd_dram 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fwd_dram.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline
wc -l gpurun_out/launches_fwd_dram.csv
