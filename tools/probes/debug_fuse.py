import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch, torch.nn as nn
from oracle import weights, vit_oracle as vo
from tpat import models_vit
T, C, B, kr = 1024, 35, 2, 0.7
sd = weights.make_audiomae_state_dict(C, T, seed=11, flavour="perturbed")
m = models_vit.vit_base_patch16(num_classes=C, drop_path_rate=0.0, mean_pooling=True, mask_2d=True, target_length=T,
                                drop_loc=(3, 6, 9), base_keep_rate=kr, precision="fp32", fuse_token=True)
m.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, 768)
m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, 768), requires_grad=False)
m.load_state_dict(sd, strict=True); m = m.cuda().eval()
x = weights.make_spectrogram("audiomae", B, T, seed=12)
with torch.no_grad():
    rl, ref = vo.forward("audiomae", sd, x, None, (3, 6, 9), kr, flag_extract_features=True, fuse_token=True)
    lg, f = m(x.cuda(), flag_extract_features=True)
for k in sorted(ref, key=lambda s: (int(s.split('.')[0].split('-')[1]), s)):
    a, b = f[k].double(), ref[k].double()
    if k.endswith("attn_score"):
        e = (a - b).abs()
        print(k, tuple(a.shape), "max rel", (e.max() / b.abs().max()).item(), "argmax col", e.max(0).values.argmax().item(), "last col err", e[:, -1].max().item(), "last col ref", b[:, -1].tolist())
    else:
        print(k, "sets equal:", all(set(p) == set(q) for p, q in zip(a.long().tolist(), b.long().tolist())))
print("---- which clip / value at col 177, block 4")
a, b = f["block-4.attn_score"].double(), ref["block-4.attn_score"].double()
print("fp32 got", a[:, 175:180].tolist(), "ref", b[:, 175:180].tolist())
print("topk pos 177 (token kept 178th) idx:", f["block-3.topk_idx"][:, 175:180].tolist(), ref["block-3.topk_idx"][:, 175:180].tolist())
sc = ref["block-3.attn_score"].double()
srt = torch.sort(sc, descending=True)
print("ref sorted score gaps around 177:", (srt.values[:, 174:181]).tolist())
m.precision = "bf16"
with torch.no_grad():
    lg2, f2 = m(x.cuda(), flag_extract_features=True)
a2 = f2["block-4.attn_score"].double()
print("bf16 got", a2[:, 175:180].tolist())
