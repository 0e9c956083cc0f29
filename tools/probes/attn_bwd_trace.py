#!/usr/bin/env python
"""Debug: build libtpat with -DTPAT_ATTN_BWD_TRACE into /tmp, run the attention backward once and print the clock stamps of
one softmax thread (slots 0 start, 1 wait S/dP, 2 S/dP ready, 3 P/dS in registers, 4 wait dq_full, 5 dq_full, 6 stored +
arrived, 7 dQ epilogue done, 8 end) merged with the MMA thread's (20 wait pds_full, 21 pds_full, 22 MMAs issued)."""
import ctypes, os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "token-pruning-audio-transformer_b200", "csrc")
out = "/tmp/libtpat_bwd_trace.so"
srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-DTPAT_ATTN_BWD_TRACE", "-o", out] + srcs)
os.environ["TPAT_LIB_PATH"] = out
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
B, N, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 513, 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).to(torch.bfloat16)
d_out = torch.randn(B * N, H * 64, device="cuda").to(torch.bfloat16)
o, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
for _ in range(3):
    ops.attention_bwd(qkv, o, d_out, lse, B, N, H, _lib.IMPL_TC)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_bwd(qkv, o, d_out, lse, B, N, H, _lib.IMPL_TC)
e1.record(); torch.cuda.synchronize()
print(f"TPAT_ATTN_BWD_WARPS={os.environ.get('TPAT_ATTN_BWD_WARPS', '16 (default)')}: {e0.elapsed_time(e1) / 10:.4f} ms per backward (traced build, back to back)")
lib = ctypes.CDLL(out)
buf = (ctypes.c_longlong * 512)()
assert lib.tpat_debug_attn_bwd_trace(buf) == 0
ev = []
for i in range(buf[254]):
    ev.append((buf[i] & ((1 << 48) - 1), buf[i] >> 48, "softmax"))
for i in range(256, buf[511]):
    ev.append((buf[i] & ((1 << 48) - 1), buf[i] >> 48, "mma"))
ev.sort()
names = {0: "start", 1: "wait s_full", 2: "s_full", 3: "P/dS in regs", 4: "wait dq_full", 5: "dq_full", 6: "stored+arrived", 7: "dQ epilogue done", 8: "end", 9: "loop top (prefetch issued)", 10: "compute entry",
         20: "wait pds_full", 21: "pds_full", 22: "MMAs issued"}
t0, prev = ev[0][0], ev[0][0]
for t, slot, who in ev:
    print(f"{t - t0:8d} (+{t - prev:6d})  {who:8s} {names.get(slot, slot)}")
    prev = t
