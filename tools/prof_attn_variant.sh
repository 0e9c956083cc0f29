#!/bin/bash
# usage: prof_attn_variant.sh <tag> [lib]   -> gpurun_out/attn_<tag>.ncu-rep
tag=$1; lib=$2
export TPAT_LIB_PATH=$lib
python tools/prof_kernels.py attn 1 > gpurun_out/pa_$tag.log 2>&1 && ncu --set full --clock-control none -k regex:"attention_tc" -c 2 -o gpurun_out/attn_$tag python tools/prof_kernels.py attn 1 >> gpurun_out/pa_$tag.log 2>&1
tail -1 gpurun_out/pa_$tag.log
