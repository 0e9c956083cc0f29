mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? =="; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run host 300 python tools/host_overhead_probe.py
TPAT_PDL=1 run host_pdl 300 python tools/host_overhead_probe.py
run ncu 1500 ncu --set full --clock-control none --import-source on -k regex:"attention|gemm_|row_bwd|colsum|layernorm|score_topk|patchify|dq_convert|attn_delta|partials" --launch-skip 20 -c 20 -f -o gpurun_out/r02_kernels python tools/profile_kernels_once.py
ls -la gpurun_out/*.ncu-rep
