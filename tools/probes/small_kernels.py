#!/usr/bin/env python
"""Device time of the small kernels around the block loop (patchify, patch-embed GEMM, pooled norm, head, top-k) at the
headline shape: 20 launches captured in a CUDA graph (no host overhead), replay time / 20, L2 warm."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import _lib
from tpat._lib import lib, check
dev = torch.device("cuda:0"); bf = torch.bfloat16
def t(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
st = lambda: torch.cuda.current_stream().cuda_stream
B, T, F, D, C = 64, 1024, 128, 768, 527
spec = torch.randn(B, T, F, device=dev)
x = torch.empty(B, 513, D, device=dev)
extra = torch.randn(1, D, device=dev); pos = torch.randn(513, D, device=dev)
pw = (torch.randn(D, 256, device=dev) * .02).to(bf); pb = torch.zeros(D, device=dev)
patches = torch.empty(B * 512, 256, device=dev, dtype=bf)
print("patchify %.1f us" % t(lambda: check(lib.tpat_patchify(spec.data_ptr(), patches.data_ptr(), _lib.BF16, x.data_ptr(), extra.data_ptr(), pos.data_ptr(), B, T, F, D, 1, _lib.TOKENS_TIME_MAJOR, st()), "p")))
print("patch gemm %.1f us" % t(lambda: check(lib.tpat_gemm(patches.data_ptr(), _lib.BF16, 256, pw.data_ptr(), _lib.BF16, pb.data_ptr(), x.data_ptr(), _lib.F32, D, None, 0, pos.data_ptr(), 512, 1, B * 512, D, 256, _lib.EPI_BIAS_POS, _lib.IMPL_TC, st()), "g")))
x2 = torch.randn(B, 178, D, device=dev); g1 = torch.ones(D, device=dev); b0 = torch.zeros(D, device=dev)
pooled = torch.empty(B, D, device=dev)
print("pool_norm %.1f us" % t(lambda: check(lib.tpat_pool_norm(x2.data_ptr(), pooled.data_ptr(), g1.data_ptr(), b0.data_ptr(), 1e-6, None, None, 0.0, B, 178, D, _lib.VARIANT_AUDIOMAE, st()), "pn")))
hw = torch.randn(C, D, device=dev) * .02; hb = torch.zeros(C, device=dev); logits = torch.empty(B, C, device=dev)
print("head %.1f us" % t(lambda: check(lib.tpat_head(pooled.data_ptr(), hw.data_ptr(), hb.data_ptr(), logits.data_ptr(), B, D, C, st()), "h")))
ref = pooled @ hw.t() + hb
print("head max err %.2e" % (logits - ref).abs().max().item())
part = torch.rand(B, 60, 513, device=dev); score = torch.empty(B, 512, device=dev); idx = torch.empty(B, 359, device=dev, dtype=torch.int64)
print("score_topk(R=60,N=513,k=359) %.1f us" % t(lambda: check(lib.tpat_score_topk(part.data_ptr(), 60, 6144.0, score.data_ptr(), idx.data_ptr(), None, B, 513, 1, 359, st()), "t")))
