mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_train.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_train.log; wc -l gpurun_out/launches_train.csv
