#!/bin/bash
# Build an experimental libtpat variant: tools/build_variant.sh <out.so> [extra nvcc flags...]
out=$1; shift
cd "$(dirname "$0")/../token-pruning-audio-transformer_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared "$@" -o "$out" *.cu
