import os, sys
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat import ops, _lib
dev=torch.device("cuda:0"); bf=torch.bfloat16
M,N,K=64*513,3072,768
a=torch.randn(M,K,device=dev).to(bf); w=(torch.randn(N,K,device=dev)*0.04).to(bf); b=torch.randn(N,device=dev)*0.5
c=torch.empty(M,N,device=dev,dtype=bf)
for _ in range(3): ops.gemm(a,w,b,bf,_lib.EPI_BIAS_GELU,_lib.IMPL_TC,out=c)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): ops.gemm(a,w,b,bf,_lib.EPI_BIAS_GELU,_lib.IMPL_TC,out=c)
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/20
pre=(a[:4096].float()@w.float().t()+b)
ref=torch.nn.functional.gelu(pre.double()).float()
got=c[:4096].float()
err=(got-ref).abs()
refb=ref.to(bf).float()
print(os.environ.get("TPAT_LIB_PATH","default"), f"fc1 {ms:.4f} ms {2*M*N*K/ms/1e9:.0f} TF/s; max abs err {err.max():.3e}; mean abs err {err.mean():.3e}; bf16-rounding-only mean err {(refb-ref).abs().mean():.3e}; pre std {pre.std():.2f}")
for lo, hi in ((-9, -4), (-4, -3), (-3, -2), (-2, -1), (-1, 0), (0, 2), (2, 9)):
    m = (pre >= lo) & (pre < hi)
    if m.any():
        print(f"  x in [{lo},{hi}): n={int(m.sum()):8d} max abs err {err[m].max():.2e} mean {err[m].mean():.2e}  (bf16-only mean {(refb-ref).abs()[m].mean():.2e}, |gelu| mean {ref[m].abs().mean():.2e})")
