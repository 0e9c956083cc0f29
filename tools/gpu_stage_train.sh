mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t12a 120 python -m pytest tests/test_gpu_12_attention_v3.py -q -m gpu -s -x -k "matches_reference and 66"
run t12 240 python -m pytest tests/test_gpu_12_attention_v3.py -q -m gpu -s
run v3b 200 python tools/attn_v3_bench.py
