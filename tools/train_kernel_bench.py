#!/usr/bin/env python
"""CUDA-event timings of the fine-tune step's kernels at the benchmark shapes (B = 64, N = 513 / 360 / 253 / 178),
each kernel alone, L2 flushed between launches.  Output: one line per kernel with its roofline figure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


B, H, D, Dh = 64, 12, 768, 3072
bf = torch.bfloat16
for N in (513, 360, 253, 178):
    M = B * N
    qkv = (torch.randn(M, 3 * D, device=dev) * 0.5).to(bf)
    d_out = torch.randn(M, D, device=dev).to(bf)
    out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    t_f = timeit(lambda: ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC))
    t_b = timeit(lambda: ops.attention_bwd(qkv, out, d_out, lse, B, N, H, _lib.IMPL_TC))
    fl = 4.0 * B * N * N * D
    print(f"N={N:4d} attention fwd {t_f:.4f} ms ({fl / t_f / 1e9:7.1f} TF/s)   bwd (delta + memset + kernel + convert) {t_b:.4f} ms "
          f"({2.5 * fl / t_b / 1e9:7.1f} TF/s)")
    dy = (torch.randn(M, Dh, device=dev) * 0.1).to(bf); x = torch.randn(M, D, device=dev).to(bf)
    dw = torch.zeros(Dh, D, device=dev)
    t = timeit(lambda: ops.gemm_wgrad(dy, x, out=dw))
    print(f"        wgrad fc1 [{Dh}x{D}] K={M}: {t:.4f} ms ({2.0 * M * Dh * D / t / 1e9:7.1f} TF/s)")
    dy2 = dy[:, :D].contiguous(); dw2 = torch.zeros(D, D, device=dev)
    t = timeit(lambda: ops.gemm_wgrad(dy2, x, out=dw2))
    print(f"        wgrad proj [{D}x{D}] K={M}: {t:.4f} ms ({2.0 * M * D * D / t / 1e9:7.1f} TF/s)")
    g = torch.randn(M, D, device=dev).to(bf); w2t = (torch.randn(Dh, D, device=dev) * 0.02).to(bf)
    pre = torch.randn(M, Dh, device=dev).to(bf)
    t = timeit(lambda: ops.gemm_train(g, w2t, None, bf, _lib.EPI_DGELU, _lib.IMPL_TC, aux=pre))
    print(f"        dgrad fc2 + GELU' [M x {Dh}] K={D}: {t:.4f} ms ({2.0 * M * Dh * D / t / 1e9:7.1f} TF/s)")
    w1t = (torch.randn(D, Dh, device=dev) * 0.02).to(bf)
    t = timeit(lambda: ops.gemm_train(dy, w1t, None, torch.float32, _lib.EPI_BIAS, _lib.IMPL_TC))
    print(f"        dgrad fc1 (fp32 out) [M x {D}] K={Dh}: {t:.4f} ms ({2.0 * M * Dh * D / t / 1e9:7.1f} TF/s)")
    t = timeit(lambda: ops.colsum(dy))
    print(f"        colsum [M x {Dh}] bf16: {t:.4f} ms ({M * Dh * 2 / t / 1e6:7.1f} GB/s)")
    xs = torch.randn(B, N, D, device=dev); dyf = torch.randn(B, N, D, device=dev); gu = torch.randn(B, N, D, device=dev)
    gam = torch.ones(D, device=dev)
    t = timeit(lambda: ops.row_bwd(dyf, xs, gam, gu, bf))
    print(f"        row_bwd (LN backward + residual + bf16 copy + 3 column sums): {t:.4f} ms ({M * D * (12 + 6) / t / 1e6:7.1f} GB/s)")
