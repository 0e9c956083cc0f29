import os, sys
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
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
for N in (512, 513, 514, 384, 385, 640, 641):
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).to(torch.bfloat16)
    a = t(lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC))
    b = t(lambda: ops.attention(qkv, B, N, H, 1, _lib.SCORE_COLMEAN, _lib.IMPL_TC))
    print(f"N={N}: {a:.4f} / {b:.4f}")
