#!/usr/bin/env python
"""A/B of the single-pass attention kernels at the benchmark shapes, L2 flushed between launches:
old = attention_tc.cu<false>, v3 = attention_tc3.cu (TPAT_ATTN_V3=1), v4 = attention_tc4.cu (TPAT_ATTN_V4=1, default),
v5 = attention_tc5.cu (TPAT_ATTN_V5=1).
Variants are interleaved per repetition so that clock / power drift hits all of them alike."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, H = 64, 12
VARIANTS = {"old": {"TPAT_ATTN_V3": "0", "TPAT_ATTN_V4": "0", "TPAT_ATTN_V5": "0"}, "v3": {"TPAT_ATTN_V3": "1", "TPAT_ATTN_V4": "0", "TPAT_ATTN_V5": "0"},
            "v4": {"TPAT_ATTN_V3": "0", "TPAT_ATTN_V4": "1", "TPAT_ATTN_V5": "0"}, "v5": {"TPAT_ATTN_V3": "0", "TPAT_ATTN_V4": "1", "TPAT_ATTN_V5": "1"}}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["old", "v4"]


def once(fn):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for N in (513, 514, 512, 360, 253, 178):
    qkv = (torch.randn(B * N, 3 * H * 64, device=dev) * 1.0).to(torch.bfloat16)
    ts = {v: [] for v in which}
    fn = lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    for rep in range(13):
        for v in which:
            os.environ.update(VARIANTS[v])
            t = once(fn)
            if rep >= 3:
                ts[v].append(t)
    fl = 4.0 * B * N * N * 768
    med = {v: sorted(x)[len(x) // 2] for v, x in ts.items()}
    print(f"N={N:4d}: " + "   ".join(f"{v} {med[v]:.4f} ms ({fl / med[v] / 1e9:6.1f} TF/s)" for v in which) +
          (f"   old/v4 x{med['old'] / med['v4']:.3f}" if "old" in med and "v4" in med else "") +
          (f"   v4/v5 x{med['v4'] / med['v5']:.3f}" if "v4" in med and "v5" in med else ""))
