#!/usr/bin/env python
"""Throughput sweeps for BASELINE.json configs[2] (keep-rate sweep vs unpruned) and configs[4] (batch sweep),
AudioMAE and AST ViT-B/16 1024x128, bf16 tensor-core path, one B200.  Writes a markdown table to stdout.
    python tools/sweep.py [--graph]
"""
import argparse, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch, torch.nn as nn
from oracle import weights
from tpat import models_vit, ASTModel

ap = argparse.ArgumentParser(); ap.add_argument("--graph", action="store_true"); args = ap.parse_args()
dev = torch.device("cuda:0")
T = 1024


def flops(n_patches, extra, kr, drop_loc=(3, 6, 9), D=768, Dh=3072, C=527):
    fl = 2.0 * n_patches * 256 * D; cur = n_patches
    for i in range(12):
        n_in = cur + extra
        if i in drop_loc and kr < 1.0: cur = math.ceil(kr * cur)
        n_out = cur + extra
        fl += 2.0 * n_in * D * 3 * D + 4.0 * n_in * n_in * D + 2.0 * n_in * D * D + 4.0 * n_out * D * Dh
    return fl + 2.0 * D * C


def build(variant, kr):
    if variant == "audiomae":
        m = models_vit.vit_base_patch16(num_classes=527, drop_path_rate=0.1, mean_pooling=True, mask_2d=True, target_length=T,
                                        drop_loc=(3, 6, 9), base_keep_rate=kr, precision="bf16")
        m.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, 768)
        m.pos_embed = nn.Parameter(torch.zeros(1, 513, 768), requires_grad=False)
        m.load_state_dict(weights.make_audiomae_state_dict(527, T, 0), strict=True)
    else:
        m = ASTModel(label_dim=527, input_tdim=T, imagenet_pretrain=False, audioset_pretrain=False, verbose=False,
                     drop_loc=(3, 6, 9), base_keep_rate=kr, precision="bf16")
        m.load_state_dict(weights.make_ast_state_dict(527, T, 0), strict=False)
    m = m.to(dev).eval(); m.use_cuda_graph = args.graph
    return m


def timeit(m, x, steps):
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): m(x)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print("### keep-rate sweep, B = 64 (configs[2]: pruned vs unpruned; FLOP column = analytic post-pruning count)\n")
print("| model | keep | GFLOP/clip | ms / 64 clips | clips/s | speed-up vs unpruned | FLOP ratio unpruned/pruned | TFLOP/s |")
print("|---|---|---|---|---|---|---|---|")
for variant, extra in (("audiomae", 1), ("ast", 2)):
    base = None
    for kr in (1.0, 0.9, 0.8, 0.7, 0.6, 0.5):
        m = build(variant, kr)
        x = weights.make_spectrogram(variant, 64, T).to(dev)
        ms = timeit(m, x, 20)
        fl = flops(512, extra, kr)
        base = base or (ms, fl)
        print(f"| {variant} | {kr} | {fl / 1e9:.2f} | {ms:.3f} | {64e3 / ms:.0f} | {base[0] / ms:.3f} | {base[1] / fl:.3f} | {64 * fl / ms / 1e9:.0f} |", flush=True)
        del m
print("\n### batch sweep, keep 0.7 vs unpruned (configs[4])\n")
print("| model | batch | pruned ms | pruned clips/s | unpruned ms | unpruned clips/s | speed-up |")
print("|---|---|---|---|---|---|---|")
for variant in ("audiomae", "ast"):
    mp, mu = build(variant, 0.7), build(variant, 1.0)
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        x = weights.make_spectrogram(variant, B, T).to(dev)
        steps = 30 if B <= 64 else 8
        a, b = timeit(mp, x, steps), timeit(mu, x, steps)
        print(f"| {variant} | {B} | {a:.3f} | {B * 1e3 / a:.0f} | {b:.3f} | {B * 1e3 / b:.0f} | {b / a:.3f} |", flush=True)
    del mp, mu
