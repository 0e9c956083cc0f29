#!/usr/bin/env python
"""GEMM timing probe: TF/s of the tcgen05 GEMM for a few shapes (optionally under TPAT_GEMM_DEBUG_SKIP)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (M, N, K, epi, od) in [(32832, 2304, 768, _lib.EPI_BIAS, bf), (32832, 3072, 768, _lib.EPI_BIAS, bf), (32832, 3072, 768, _lib.EPI_BIAS_GELU, bf),
                           (32832, 768, 3072, _lib.EPI_BIAS, bf), (32768, 4096, 4096, _lib.EPI_BIAS, bf), (8192, 8192, 8192, _lib.EPI_BIAS, bf)]:
    a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) * .02).to(bf); b = torch.zeros(N, device=dev)
    c = torch.empty(M, N, device=dev, dtype=od)
    ms = t(lambda: ops.gemm(a, w, b, od, epi, _lib.IMPL_TC, out=c))
    print(f"M={M} N={N} K={K} epi={epi}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.0f} TF/s", flush=True)
    if M == 8192:
        ms = t(lambda: torch.matmul(a, w.T))
        print(f"   torch.matmul (cuBLAS) same shape: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.0f} TF/s")
