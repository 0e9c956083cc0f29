#!/usr/bin/env python
"""How long does the HOST take to enqueue one fine-tune step (no synchronisation inside the loop) against the GPU time of
the step?  If the two are close the step is launch / Python bound, not kernel bound."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
import torch.nn.functional as F
import bench
from tpat.lr_decay import param_groups_lrd
from tpat.optim import FusedAdamW

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
opt = FusedAdamW(param_groups_lrd(model, 0.05, model.no_weight_decay(), 0.75), lr=1e-3, betas=(0.9, 0.95), model=model)
x = torch.randn(64, 1, 1024, 128, device=dev) * 0.5
y = (torch.rand(64, 527, device=dev) < 0.01).float()


def step(parts):
    t0 = time.perf_counter()
    logits = model(x)
    t1 = time.perf_counter()
    loss = F.binary_cross_entropy_with_logits(logits, y)
    opt.zero_grad()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    for k, v in zip(("fwd", "loss+zero", "bwd", "opt"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
        parts[k] = parts.get(k, 0.0) + v


for _ in range(3):
    step({})
torch.cuda.synchronize()
n = 10
parts = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    step(parts)
e1.record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host enqueue {host / n * 1e3:.2f} ms/step, GPU {e0.elapsed_time(e1) / n:.2f} ms/step; host parts (ms): "
      + ", ".join(f"{k} {v / n * 1e3:.2f}" for k, v in parts.items()))
