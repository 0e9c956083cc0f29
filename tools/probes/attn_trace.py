#!/usr/bin/env python
"""Debug: build libtpat with -DTPAT_ATTN_TRACE into /tmp, run one attention launch and print the clock stamps of one
softmax thread (phases: 2 wait S, 3 S ready, 4 S in registers, 5 max/rescale done, 6 P buffer free, 7 exp+stores done,
8 arrived, 9 loop done, 10 O ready, 11 end)."""
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
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.0).to(torch.bfloat16)
o = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
partial = torch.empty(B, H * 8, N, device="cuda")
lib.tpat_attention.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
for _ in range(3):
    rc = lib.tpat_attention(qkv.data_ptr(), o.data_ptr(), 1, partial.data_ptr(), mode, B, N, H, 64, 1, 0.125, 1, None)
    assert rc == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
assert lib.tpat_debug_attn_trace(buf) == 0
n = buf[127]
prev = 0
names = {1: "start", 2: "wait S", 3: "S ready", 4: "S in regs", 5: "max done", 6: "P free", 7: "exp+store done", 8: "arrived", 9: "loop done", 10: "O ready", 11: "end", 12: "exps+st issued", 13: "max/vote/flag", 14: "pair barrier"}
for i in range(n):
    slot, t = buf[i] >> 48, buf[i] & ((1 << 48) - 1)
    print(f"{t:8d} (+{t - prev:6d})  {names.get(slot, slot)}")
    prev = t
