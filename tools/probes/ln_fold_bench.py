#!/usr/bin/env python
"""Per-kernel cost of the LayerNorm fold (tpat_gemm_ln) at the headline shape: producer (proj / fc2 + emission) and consumer (qkv / fc1 folded)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps * 1e3
M, D, Dh = 64 * 513, 768, 3072
x = torch.randn(M, D, device=dev)
ao = torch.randn(M, D, device=dev).to(bf); hid = torch.randn(M, Dh, device=dev).to(bf)
wp = (torch.randn(D, D, device=dev) * .02).to(bf); w2 = (torch.randn(D, Dh, device=dev) * .02).to(bf); b = torch.zeros(D, device=dev)
print("proj plain %.1f us" % t(lambda: ops.gemm(ao, wp, b, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)))
print("proj emit  %.1f us" % t(lambda: ops.gemm_ln(ao, wp, b, torch.float32, _lib.EPI_BIAS_RESIDUAL, residual=x, out=x, emit=True)))
print("fc2 plain  %.1f us" % t(lambda: ops.gemm(hid, w2, b, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)))
print("fc2 emit   %.1f us" % t(lambda: ops.gemm_ln(hid, w2, b, torch.float32, _lib.EPI_BIAS_RESIDUAL, residual=x, out=x, emit=True)))
_, xb, part = ops.gemm_ln(ao, wp, b, torch.float32, _lib.EPI_BIAS_RESIDUAL, residual=x, out=x, emit=True)
wq = (torch.randn(3 * D, D, device=dev) * .02).to(bf); bq = torch.zeros(3 * D, device=dev); cs = wq.float().sum(1).contiguous()
w1 = (torch.randn(Dh, D, device=dev) * .02).to(bf); b1 = torch.zeros(Dh, device=dev); cs1 = w1.float().sum(1).contiguous()
oq = torch.empty(M, 3 * D, device=dev, dtype=bf); oh = torch.empty(M, Dh, device=dev, dtype=bf)
print("qkv plain  %.1f us" % t(lambda: ops.gemm(xb, wq, bq, bf, _lib.EPI_BIAS, _lib.IMPL_TC, out=oq)))
print("qkv fold   %.1f us" % t(lambda: ops.gemm_ln(xb, wq, bq, bf, _lib.EPI_BIAS, out=oq, ln_part=part, ln_colsum=cs)))
print("fc1 plain  %.1f us" % t(lambda: ops.gemm(xb, w1, b1, bf, _lib.EPI_BIAS_GELU, _lib.IMPL_TC, out=oh)))
print("fc1 fold   %.1f us" % t(lambda: ops.gemm_ln(xb, w1, b1, bf, _lib.EPI_BIAS_GELU, out=oh, ln_part=part, ln_colsum=cs1)))
