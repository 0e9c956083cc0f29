#!/usr/bin/env python
"""Attention-only timing at the four block sizes (uses TPAT_LIB_PATH if set)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0"); B, H = 64, 12
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
out = []
for N in (513, 360, 253, 178):
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).to(torch.bfloat16)
    a = t(lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC))
    b = t(lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_COLMEAN, _lib.IMPL_TC))
    out.append(f"N={N}: {a:.4f} / {b:.4f}")
print(os.environ.get("TPAT_LIB_PATH", "default"), " | ".join(out), " (ms none / colmean)")
