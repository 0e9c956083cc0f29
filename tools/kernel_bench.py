#!/usr/bin/env python
"""Per-kernel CUDA-event timings at the headline shapes (AudioMAE 1024x128, B=64, keep 0.7).
Usage: python tools/kernel_bench.py [--batch 64] [--reps 20]   (GPU box only)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch  # noqa: E402
from tpat import ops, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--extra", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda:0")
B, D, H, Dh = args.batch, 768, 12, 3072
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=args.reps, flush_l2=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


rows = []
bf = torch.bfloat16
sched = [(513, 4, 1), (360, 3, 1), (253, 3, 1), (178, 2, 0)]   # (N, blocks at this N, pruning blocks among them)
total_ms = 0.0
for N, nblk, nprune in sched:
    M = B * N
    x = torch.randn(M, D, device=dev)
    g = torch.ones(D, device=dev); b0 = torch.zeros(D, device=dev)
    y = torch.randn(M, D, device=dev).to(bf)
    wqkv = (torch.randn(3 * D, D, device=dev) * .02).to(bf); bqkv = torch.zeros(3 * D, device=dev)
    wproj = (torch.randn(D, D, device=dev) * .02).to(bf); bproj = torch.zeros(D, device=dev)
    w1 = (torch.randn(Dh, D, device=dev) * .02).to(bf); b1 = torch.zeros(Dh, device=dev)
    w2 = (torch.randn(D, Dh, device=dev) * .02).to(bf); b2 = torch.zeros(D, device=dev)
    qkv = torch.randn(M, 3 * D, device=dev).to(bf)
    hid = torch.randn(M, Dh, device=dev).to(bf)
    oq = torch.empty(M, 3 * D, device=dev, dtype=bf); oh = torch.empty(M, Dh, device=dev, dtype=bf)
    res = {}
    res["ln"] = (timeit(lambda: ops.layernorm(x, g, b0, 1e-6, bf)), M * D * 6 / 1e9, "GB")
    res["qkv"] = (timeit(lambda: ops.gemm(y, wqkv, bqkv, bf, _lib.EPI_BIAS, _lib.IMPL_TC, out=oq)), 2.0 * M * 3 * D * D / 1e12, "TF")
    res["attn"] = (timeit(lambda: ops.attention(qkv, B, N, H, args.extra, _lib.SCORE_NONE, _lib.IMPL_TC)), 4.0 * B * N * N * D / 1e12, "TF")
    res["attn_score"] = (timeit(lambda: ops.attention(qkv, B, N, H, args.extra, _lib.SCORE_COLMEAN, _lib.IMPL_TC)), 4.0 * B * N * N * D / 1e12, "TF")
    res["proj"] = (timeit(lambda: ops.gemm(y, wproj, bproj, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)), 2.0 * M * D * D / 1e12, "TF")
    res["fc1"] = (timeit(lambda: ops.gemm(y, w1, b1, bf, _lib.EPI_BIAS_GELU, _lib.IMPL_TC, out=oh)), 2.0 * M * Dh * D / 1e12, "TF")
    res["fc1_nogelu"] = (timeit(lambda: ops.gemm(y, w1, b1, bf, _lib.EPI_BIAS, _lib.IMPL_TC, out=oh)), 2.0 * M * Dh * D / 1e12, "TF")
    res["fc2"] = (timeit(lambda: ops.gemm(hid, w2, b2, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)), 2.0 * M * D * Dh / 1e12, "TF")
    partial = torch.rand(B, H * ops.attention_qtiles(N, _lib.IMPL_TC), N, device=dev)
    k = int(0.7 * (N - 1)) + 1
    res["topk"] = (timeit(lambda: ops.score_topk(partial, 1.0, 1, k)), partial.numel() * 4 / 1e9, "GB")
    idx = torch.stack([torch.randperm(N - 1, device=dev)[:k] for _ in range(B)])
    x3 = x.reshape(B, N, D)
    res["gather_ln"] = (timeit(lambda: ops.gather_layernorm(x3, idx, 1, g, b0, 1e-6, bf)), B * (k + 1) * D * 10 / 1e9, "GB")
    blk = 2 * res["ln"][0] + res["qkv"][0] + res["attn"][0] + res["proj"][0] + res["fc1"][0] + res["fc2"][0]
    total_ms += nblk * blk + nprune * (res["attn_score"][0] - res["attn"][0] + res["topk"][0])
    for name, (ms, work, unit) in res.items():
        rate = work / (ms * 1e-3)
        rows.append(dict(N=N, kernel=name, ms=round(ms, 4), rate=round(rate, 1), unit=unit + "/s"))
        print(f"N={N:4d} {name:12s} {ms:8.4f} ms  {rate:9.1f} {unit}/s")
    print(f"N={N:4d} block total (no prune) {blk:.3f} ms")
print(f"estimated forward (sum of kernels, cold L2): {total_ms:.3f} ms -> {B / total_ms * 1e3:.0f} clips/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)
