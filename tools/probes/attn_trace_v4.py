#!/usr/bin/env python
"""Debug: build libtpat with -DTPAT_ATTN_TRACE into /tmp, run one v4 attention launch and print the clock stamps of one
softmax thread (warp 2, lane 0: quarter 2, half 0) and of the MMA thread of the same CTA, on a common time axis.
softmax: 2 wait S, 3 S ready, 4 S in registers, 5 exps + P store issued, 6 P in TMEM, 7 arrived, 9 loop done, 10 O ready,
11 end.  MMA: 20 Q ready, 21 S(j+1) issued, 22 P_j ready (P.V issue follows)."""
import ctypes, os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "token-pruning-audio-transformer_b200", "csrc")
out = "/tmp/libtpat_trace.so"
srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-DTPAT_ATTN_TRACE", "-o", out] + srcs)
import torch
lib = ctypes.CDLL(out)
B, N, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 513, 12
os.environ["TPAT_ATTN_V4"] = "1"
os.environ["TPAT_ATTN_V5"] = sys.argv[2] if len(sys.argv) > 2 else "0"     # 1: the persistent kernel (CTA 100, its first items)
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.0).to(torch.bfloat16)
o = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
partial = torch.empty(B, H * 8, N, device="cuda")
lib.tpat_attention.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
for _ in range(3):
    rc = lib.tpat_attention(qkv.data_ptr(), o.data_ptr(), 1, partial.data_ptr(), 0, B, N, H, 64, 1, 0.125, 1, None)
    assert rc == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
assert lib.tpat_debug_attn_trace(buf) == 0
names = {1: "start", 2: "wait S", 3: "S ready", 4: "S in regs", 5: "exps+st issued", 6: "P in TMEM", 7: "arrived", 9: "loop done",
         10: "O ready", 11: "end / store issued", 12: "O in regs, o_empty arrived", 13: "staged + CTA barrier", 20: "MMA: Q ready",
         21: "MMA: S(j+1) issued", 22: "MMA: P_j ready", 23: "MMA: o_empty ready"}
t0_s, t0_m = buf[126], buf[254]
ev = []
for base, n, t0 in ((0, buf[127], t0_s), (128, buf[255], t0_m)):
    for i in range(n):
        slot, t = buf[base + i] >> 48, buf[base + i] & ((1 << 48) - 1)
        ev.append((t + (t0 - t0_s), slot))
ev.sort()
prev = {0: 0, 1: 0}
for t, slot in ev:
    who = 1 if slot >= 20 else 0
    print(f"{t:8d} (+{t - prev[who]:6d})  {'        ' if who else ''}{names.get(slot, slot)}")
    prev[who] = t
