#!/usr/bin/env python
"""Latency at small batch (CUDA-graph replay): AudioMAE 1024x128 keep 0.7, B in 1..16."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch, torch.nn as nn
from tpat import models_vit
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = models_vit.vit_base_patch16(num_classes=527, drop_path_rate=0.1, mean_pooling=True, mask_2d=True, target_length=1024,
                                drop_loc=(3, 6, 9), base_keep_rate=0.7, precision="bf16")
m.patch_embed = models_vit.PatchEmbed((1024, 128), 16, 1, 768)
m.pos_embed = nn.Parameter(torch.zeros(1, 513, 768), requires_grad=False)
m = m.to(dev).eval(); m.use_cuda_graph = True
out = []
with torch.no_grad():
    for B in (1, 2, 4, 8, 16):
        x = torch.randn(B, 1, 1024, 128, device=dev) * 0.5
        for _ in range(5): m(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50): m(x)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 50 * 1e3
        out.append(f"B={B}: {ms:.3f} ms")
print(os.environ.get("TPAT_GEMM_2CTA", "2cta"), " | ".join(out))
