#!/usr/bin/env python
"""Small launches of the attention kernels (old single-pass, attention_tc4, attention_tc5, two-pass + column mean, AST cls row)
at tile / key-block tail shapes and with several items per CTA: a quick all-variants smoke (also usable under compute-sanitizer where the pool allows it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0")
H = 2
for name, env in (("old", {"TPAT_ATTN_V4": "0", "TPAT_ATTN_V5": "0"}), ("v4", {"TPAT_ATTN_V4": "1", "TPAT_ATTN_V5": "0"}),
                  ("v5", {"TPAT_ATTN_V4": "1", "TPAT_ATTN_V5": "1"})):
    os.environ.update(env)
    for B, N in ((2, 2), (2, 33), (3, 65), (2, 129), (2, 200), (80, 130)):
        qkv = torch.randn(B * N, 3 * H * 64, device=dev).to(torch.bfloat16)
        out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
        ops.attention(qkv, B, N, H, 1, _lib.SCORE_COLMEAN, _lib.IMPL_TC)
        if N > 2:
            ops.attention(qkv, B, N, H, 2, _lib.SCORE_CLS_ROW, _lib.IMPL_TC)
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    print(name, "ok", flush=True)
