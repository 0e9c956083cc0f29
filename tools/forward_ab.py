#!/usr/bin/env python
"""Whole-forward A/B on one B200: the benchmark workload (AudioMAE ViT-B/16 1024x128, keep 0.7, 64 clips, bf16, CUDA-graph
replay on rotating input batches) timed under the environment it is started with.  One process per variant, e.g.
    for v in "" "TPAT_ATTN_V4=1" "TPAT_L2_PERSIST_MB=64"; do env $v python tools/forward_ab.py "$v"; done
Prints: label, ms / step (median of `reps` bursts of `steps` replays), clips/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch, torch.nn as nn
from oracle import weights
from tpat import models_vit, ASTModel

label = sys.argv[1] if len(sys.argv) > 1 else "default"
variant = os.environ.get("AB_VARIANT", "audiomae")
kr = float(os.environ.get("AB_KEEP", "0.7"))
steps, reps, B, T = int(os.environ.get("AB_STEPS", "20")), int(os.environ.get("AB_REPS", "7")), 64, 1024
dev = torch.device("cuda:0")
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
m = m.to(dev).eval(); m.use_cuda_graph = True
xs = [weights.make_spectrogram(variant, B, T, seed=1234 + i).to(dev) for i in range(4)]
with torch.no_grad():
    for x in xs:
        for _ in range(2): m(x)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps): m(xs[s % len(xs)])
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / steps)
        torch.cuda._sleep(int(2e8))      # let the power state relax between bursts
        torch.cuda.synchronize()
with torch.no_grad():
    chk = m(xs[0]).float()
import hashlib
digest = hashlib.sha256(chk.cpu().numpy().tobytes()).hexdigest()[:12]
med = sorted(ts)[len(ts) // 2]
print(f"{label or 'default':40s} {variant} kr={kr}: {med:.4f} ms/step (min {min(ts):.4f}, max {max(ts):.4f})  {B * 1e3 / med:8.0f} clips/s  logits sha {digest}", flush=True)
