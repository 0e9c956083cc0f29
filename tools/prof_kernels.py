#!/usr/bin/env python
"""Launch each hot kernel a few times at headline shapes (for `ncu --set full -k regex:...`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch  # noqa: E402
from tpat import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
B, N, D, H, Dh = 64, 513, 768, 12, 3072
M = B * N
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.randn(M, D, device=dev)
y = torch.randn(M, D, device=dev).to(bf)
g = torch.ones(D, device=dev); b0 = torch.zeros(D, device=dev)
for _ in range(reps):
    if which in ("all", "attn"):
        qkv = torch.randn(M, 3 * D, device=dev).to(bf)
        ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
        ops.attention(qkv, B, N, H, 1, _lib.SCORE_COLMEAN, _lib.IMPL_TC)
    if which in ("all", "gemm"):
        w1 = (torch.randn(Dh, D, device=dev) * .02).to(bf); b1 = torch.zeros(Dh, device=dev)
        oh = torch.empty(M, Dh, device=dev, dtype=bf)
        ops.gemm(y, w1, b1, bf, _lib.EPI_BIAS_GELU, _lib.IMPL_TC, out=oh)
        wp = (torch.randn(D, D, device=dev) * .02).to(bf)
        ops.gemm(y, wp, b0, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)
        wq = (torch.randn(3 * D, D, device=dev) * .02).to(bf); bq = torch.zeros(3 * D, device=dev)
        oq = torch.empty(M, 3 * D, device=dev, dtype=bf)
        ops.gemm(y, wq, bq, bf, _lib.EPI_BIAS, _lib.IMPL_TC, out=oq)
    if which in ("all", "ln"):
        ops.layernorm(x, g, b0, 1e-6, bf)
torch.cuda.synchronize()
print("ok")
