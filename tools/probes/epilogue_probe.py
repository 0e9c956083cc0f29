#!/usr/bin/env python
"""Where does the bf16-output GEMM epilogue's time go?  Times qkv / fc1+GELU / fc1+GELU+derivative / dgrad+GELU' at
M = 64 x 513 with the library named by TPAT_LIB_PATH (debug variants: no aux loads, no output stores)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M, D, Dh = 64 * 513, 768, 3072
bf = torch.bfloat16
y = torch.randn(M, D, device=dev).to(bf)
w1 = (torch.randn(Dh, D, device=dev) * 0.02).to(bf); b1 = torch.zeros(Dh, device=dev)
wq = (torch.randn(3 * D, D, device=dev) * 0.02).to(bf); bq = torch.zeros(3 * D, device=dev)
aux = torch.rand(M, Dh, device=dev).to(bf)


def t(fn, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


h = torch.randn(M, Dh, device=dev).to(bf)
w2 = (torch.randn(D, Dh, device=dev) * 0.02).to(bf); b2 = torch.zeros(D, device=dev)
xr = torch.randn(M, D, device=dev)
print(os.environ.get("TPAT_LIB_PATH", "default"), os.environ.get("TPAT_GEMM_RES_L2_PREFETCH", "-"), os.environ.get("TPAT_GEMM_NO_L2_PREFETCH", "-"),
      "fc2+res %.4f" % t(lambda: ops.gemm(h, w2, b2, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=xr, out=xr)),
      "qkv %.4f" % t(lambda: ops.gemm(y, wq, bq, bf, _lib.EPI_BIAS, _lib.IMPL_TC)),
      "fc1+gelu %.4f" % t(lambda: ops.gemm(y, w1, b1, bf, _lib.EPI_BIAS_GELU, _lib.IMPL_TC)),
      "fc1+gelu+dact %.4f" % t(lambda: ops.gemm_train(y, w1, b1, bf, _lib.EPI_BIAS_GELU, _lib.IMPL_TC, want_dact=True)),
      "dgelu %.4f" % t(lambda: ops.gemm_train(y, w1, None, bf, _lib.EPI_DGELU, _lib.IMPL_TC, aux=aux)))
