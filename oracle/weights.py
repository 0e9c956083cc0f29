"""TEST INFRASTRUCTURE ONLY.  Deterministic state-dict generators for the parity fixtures.

The 86 M-parameter state-dict cannot be committed, so every consumer (golden generator here,
parity tests and bench on the GPU box) regenerates it from a seed with a CPU
``torch.Generator``.  ``state_dict_digest`` is stored inside each golden file so a drift of the
RNG stream is detected instead of silently comparing against the wrong weights.

Two flavours:
  * ``refinit``   -- the reference's own initialisation statistics: Linear weights
    trunc_normal(std=.02) and zero bias, LayerNorm 1/0, cls/dist/pos trunc_normal(std=.02)
    (models_vit.py:302-304,321-328; ast_models.py:326-330), conv = PyTorch Conv2d default.
    The headline bench and the BASELINE parity configs use this.
  * ``perturbed`` -- same, but biases ~ N(0, .02), LayerNorm weight 1+N(0,.1) / bias N(0,.05) so
    that every bias / affine term is exercised by the parity tests.
  * ``trained``   -- ``perturbed`` reshaped towards the statistics of a fine-tuned checkpoint: q / k weights x 3.2
    (attention logits of std ~ 3 instead of 0.3: peaked attention, importance scores that are NOT near-uniform, so
    the top-k cut is not a field of near-ties), larger proj / MLP weights, LayerNorm gains spread log-normally, and
    three outlier channels in the position table (the "massive activations" of trained ViTs).

Key names and shapes follow the reference state-dicts.
"""
import hashlib
import math

import torch


def _tn(gen, shape, std=0.02):
    t = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, std=std, generator=gen)
    return t


def _n(gen, shape, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def _block(sd, prefix, gen, dim, hidden, perturbed):
    def lin(name, out_f, in_f):
        sd[f"{prefix}{name}.weight"] = _tn(gen, (out_f, in_f))
        sd[f"{prefix}{name}.bias"] = _n(gen, (out_f,), 0.02) if perturbed else torch.zeros(out_f)

    def ln(name):
        sd[f"{prefix}{name}.weight"] = 1.0 + _n(gen, (dim,), 0.1) if perturbed else torch.ones(dim)
        sd[f"{prefix}{name}.bias"] = _n(gen, (dim,), 0.05) if perturbed else torch.zeros(dim)

    ln("norm1")
    lin("attn.qkv", 3 * dim, dim)
    lin("attn.proj", dim, dim)
    ln("norm2")
    lin("mlp.fc1", hidden, dim)
    lin("mlp.fc2", dim, hidden)


def _conv(sd, prefix, gen, dim):
    # torch.nn.Conv2d default init: kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), +1/sqrt(fan_in))
    bound = 1.0 / math.sqrt(256.0)
    sd[f"{prefix}.weight"] = (torch.rand((dim, 1, 16, 16), generator=gen) * 2 - 1) * bound
    sd[f"{prefix}.bias"] = (torch.rand((dim,), generator=gen) * 2 - 1) * bound


def _trainedlike(sd, prefix, pos_key, seed, depth, dim):
    gen = torch.Generator().manual_seed(seed + 100003)
    for i in range(depth):
        b = f"{prefix}blocks.{i}."
        sd[b + "attn.qkv.weight"][: 2 * dim] *= 3.2            # q and k rows
        sd[b + "attn.qkv.bias"][: 2 * dim] *= 3.2
        sd[b + "attn.proj.weight"] *= 2.0
        sd[b + "mlp.fc1.weight"] *= 2.0
        sd[b + "mlp.fc2.weight"] *= 1.5
        for n in ("norm1", "norm2"):
            sd[b + n + ".weight"] = sd[b + n + ".weight"] * torch.exp(0.3 * torch.randn(dim, generator=gen))
    for ch, v in ((7, 8.0), (300, -6.0), (511, 10.0)):
        sd[pos_key][..., ch] += v
    return sd


def make_audiomae_state_dict(num_classes=527, target_length=1024, seed=0, flavour="refinit",
                             depth=12, dim=768, mlp_ratio=4):
    """Keys as audiomae/models_vit.py VisionTransformer after main_finetune.py:374-382."""
    assert flavour in ("refinit", "perturbed", "trained")
    p = flavour != "refinit"
    gen = torch.Generator().manual_seed(seed)
    n_patches = (target_length // 16) * (128 // 16)
    sd = {}
    sd["cls_token"] = _tn(gen, (1, 1, dim))
    sd["pos_embed"] = _tn(gen, (1, n_patches + 1, dim))
    _conv(sd, "patch_embed.proj", gen, dim)
    for i in range(depth):
        _block(sd, f"blocks.{i}.", gen, dim, dim * mlp_ratio, p)
    sd["fc_norm.weight"] = 1.0 + _n(gen, (dim,), 0.1) if p else torch.ones(dim)
    sd["fc_norm.bias"] = _n(gen, (dim,), 0.05) if p else torch.zeros(dim)
    sd["head.weight"] = _tn(gen, (num_classes, dim))
    sd["head.bias"] = _n(gen, (num_classes,), 0.02) if p else torch.zeros(num_classes)
    if flavour == "trained":
        _trainedlike(sd, "", "pos_embed", seed, depth, dim)
    return sd


def make_ast_state_dict(label_dim=527, input_tdim=1024, seed=0, flavour="refinit",
                        depth=12, dim=768, mlp_ratio=4):
    """Keys as ast/src/models/ast_models.py ASTModel (the subset its forward uses)."""
    assert flavour in ("refinit", "perturbed", "trained")
    p = flavour != "refinit"
    gen = torch.Generator().manual_seed(seed)
    n_patches = (input_tdim // 16) * (128 // 16)
    sd = {}
    sd["v.cls_token"] = _tn(gen, (1, 1, dim))
    sd["v.dist_token"] = _tn(gen, (1, 1, dim))
    sd["v.pos_embed"] = _tn(gen, (1, n_patches + 2, dim))
    _conv(sd, "v.patch_embed.proj", gen, dim)
    for i in range(depth):
        _block(sd, f"v.blocks.{i}.", gen, dim, dim * mlp_ratio, p)
    sd["v.norm.weight"] = 1.0 + _n(gen, (dim,), 0.1) if p else torch.ones(dim)
    sd["v.norm.bias"] = _n(gen, (dim,), 0.05) if p else torch.zeros(dim)
    sd["mlp_head.0.weight"] = 1.0 + _n(gen, (dim,), 0.1) if p else torch.ones(dim)
    sd["mlp_head.0.bias"] = _n(gen, (dim,), 0.05) if p else torch.zeros(dim)
    sd["mlp_head.1.weight"] = _tn(gen, (label_dim, dim))
    sd["mlp_head.1.bias"] = _n(gen, (label_dim,), 0.02) if p else torch.zeros(label_dim)
    if flavour == "trained":
        _trainedlike(sd, "v.", "v.pos_embed", seed, depth, dim)
    return sd


def make_spectrogram(variant, batch, target_length, seed=1234):
    """x ~ N(0, 0.5^2) fp32.  AudioMAE [B,1,T,128]; AST [B,T,128]."""
    gen = torch.Generator().manual_seed(seed)
    if variant == "audiomae":
        return torch.randn(batch, 1, target_length, 128, generator=gen) * 0.5
    assert variant == "ast"
    return torch.randn(batch, target_length, 128, generator=gen) * 0.5


def state_dict_digest(sd) -> str:
    """sha256 over key names and raw fp32 bytes, in sorted key order."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()
