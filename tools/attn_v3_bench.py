#!/usr/bin/env python
"""A/B of the two single-pass attention kernels (TPAT_ATTN_V3 = 0 | 1) at the benchmark shapes, L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, H = 64, 12


def t(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for N in (513, 514, 512, 360, 253, 178):
    qkv = (torch.randn(B * N, 3 * H * 64, device=dev) * 1.0).to(torch.bfloat16)
    res = {}
    for v in ("0", "1"):
        os.environ["TPAT_ATTN_V3"] = v
        res[v] = t(lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC))
    fl = 4.0 * B * N * N * 768
    print(f"N={N:4d}: old {res['0']:.4f} ms ({fl / res['0'] / 1e9:6.1f} TF/s)   v3 {res['1']:.4f} ms ({fl / res['1'] / 1e9:6.1f} TF/s)   x{res['0'] / res['1']:.2f}")
