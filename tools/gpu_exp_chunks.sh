mkdir -p gpurun_out
: > gpurun_out/chunks_forward_ab.txt
for round in 1 2; do
  for v in "" "TPAT_MLP_CHUNKS=2" "TPAT_MLP_CHUNKS=3"; do
    env $v timeout 200 python tools/forward_ab.py "$v" >> gpurun_out/chunks_forward_ab.txt 2>> gpurun_out/chunks_forward_ab.err
  done
done
cat gpurun_out/chunks_forward_ab.txt; tail -3 gpurun_out/chunks_forward_ab.err
