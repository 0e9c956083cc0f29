#!/usr/bin/env python
"""Experiment: one batch of 64 vs two micro-batches of 32 on two CUDA streams (do the kernels of one fill the tails of the other?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch, torch.nn as nn
from oracle import weights
from tpat import models_vit
dev = torch.device("cuda:0")
def build():
    m = models_vit.vit_base_patch16(num_classes=527, drop_path_rate=0.1, mean_pooling=True, mask_2d=True, target_length=1024,
                                    drop_loc=(3, 6, 9), base_keep_rate=0.7, precision="bf16")
    m.patch_embed = models_vit.PatchEmbed((1024, 128), 16, 1, 768)
    m.pos_embed = nn.Parameter(torch.zeros(1, 513, 768), requires_grad=False)
    m.load_state_dict(weights.make_audiomae_state_dict(527, 1024, 0, "refinit"), strict=True)
    m = m.to(dev).eval(); m.use_cuda_graph = True
    return m
ma, mb, mc = build(), build(), build()
x64 = torch.randn(64, 1, 1024, 128, device=dev) * 0.5
xa, xb = x64[:32].contiguous(), x64[32:].contiguous()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
steps = 20
with torch.no_grad():
    for _ in range(3): mc(x64)
    with torch.cuda.stream(sa):
        for _ in range(3): ma(xa)
    with torch.cuda.stream(sb):
        for _ in range(3): mb(xb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps): mc(x64)
    torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(steps):
        with torch.cuda.stream(sa): ma(xa)
        with torch.cuda.stream(sb): mb(xb)
    torch.cuda.synchronize(); t2 = time.perf_counter() - t0
print(f"one stream B=64: {64 * steps / t1:.0f} clips/s ; two streams 2 x B=32: {64 * steps / t2:.0f} clips/s")
