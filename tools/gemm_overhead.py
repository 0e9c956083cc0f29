#!/usr/bin/env python
"""Fixed per-launch overhead vs per-tile time of the CTA-pair GEMM: k tiles per cluster, k = 1,2,4,8,16."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for K in (768, 3072):
    for k in (1, 2, 4, 8, 16):
        M, N = 74 * 256 * k, 256
        a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) * .02).to(bf); b = torch.zeros(N, device=dev)
        c = torch.empty(M, N, device=dev, dtype=bf)
        us = t(lambda: ops.gemm(a, w, b, bf, _lib.EPI_BIAS, _lib.IMPL_TC, out=c))
        print(f"K={K} tiles/cluster={k:2d}: {us:7.2f} us per launch, {us / k:6.2f} us per tile", flush=True)
x = torch.randn(1024, 768, device=dev); g = torch.ones(768, device=dev); bb = torch.zeros(768, device=dev)
print("tiny layernorm launch: %.2f us" % t(lambda: ops.layernorm(x, g, bb, 1e-6, bf)))
