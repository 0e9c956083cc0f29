"""Drop-in mirror of the reference's AudioMAE pruning ViT API (audiomae/models_vit.py).

Same class / factory names, constructor arguments, attribute names and state-dict keys as the
reference, so ``main_finetune.py`` / ``engine_finetune.py`` style callers
keep working:

    model = models_vit.vit_base_patch16(num_classes=527, drop_path_rate=0.1, mean_pooling=True,
                                        mask_2d=True, target_length=1024, drop_loc=(3, 6, 9),
                                        base_keep_rate=0.7)
    model.patch_embed = models_vit.PatchEmbed((1024, 128), 16, 1, 768)      # main_finetune.py:378
    model.pos_embed = nn.Parameter(torch.zeros(1, 513, 768), requires_grad=False)
    logits = model(x)                                    # x [B,1,T,128] on a B200
    logits, feats = model(x, flag_extract_features=True) # + 'mel', 'block-i.attn_score', 'block-i.topk_idx'

The sub-modules are parameter containers only.  ``forward`` never runs PyTorch math: it hands the
parameters to ``engine.ForwardEngine`` which makes one ``tpat_forward`` call into libtpat.so
(hand-written sm_100a kernels).  There is no CPU fallback.

Differences from the reference, stated (not hidden):
  * in training mode with autograd enabled ``forward`` runs the native fine-tune step (tpat/train.py: DropPath, 2-D
    token masking, activations kept, ONE autograd node whose backward is ``tpat_train_backward``); parameters then
    live in one flat buffer (``p.data`` / ``p.grad`` are views).  In eval mode / under ``no_grad`` it is the
    inference path (DropPath identity).  The ablation paths (custom_rank, drop_token_blk_idx) and
    eval-mode masking run kernel by kernel through ``ForwardEngine.run_stepwise`` (no backward).
  * torch.topk leaves the order of exactly tied scores unspecified; here ties go to the lower index.
  * extra constructor keyword ``fuse_token`` (default False): EViT's fused inattentive token, which BASELINE.json
    configs[2] names but the reference forward never implemented -- semantics from upstream EViT.
  * extra constructor keyword ``precision``: "bf16" (tcgen05 tensor-core kernels, default) or
    "fp32" (CUDA-core fp32 kernels, the index-exact parity mode).
"""
from functools import partial
from typing import Optional, Union

import torch
import torch.nn as nn

from . import _lib
from .engine import EnginePool, param_key, resolve_precision


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(t, std=0.02):
    return torch.nn.init.trunc_normal_(t, std=std)


class Mlp(nn.Module):
    """Parameter container for fc1 / fc2 (reference models_vit.py:30-46)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    """Parameter container for qkv / proj plus the pruning attributes (reference models_vit.py:49-66)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 block_id: int = 0, default_keep_rate: float = 1.0):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.num_extra_tokens = 1
        self.block_id = block_id
        self.default_keep_rate = default_keep_rate
        assert 0.0 < self.default_keep_rate <= 1.0, \
            f"default_keep_rate should be in (0, 1], got {self.default_keep_rate}"


class Block(nn.Module):
    """Parameter container for one encoder block (reference models_vit.py:138-155)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, block_id: int = 0,
                 default_keep_rate: float = 1.0):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, block_id=block_id, default_keep_rate=default_keep_rate)
        self.drop_path = nn.Identity()
        self.drop_path_rate = drop_path
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.block_id = block_id
        self.num_extra_tokens = 1


class PatchEmbed(nn.Module):
    """Parameter container for the 16x16 / stride-16 projection (reference models_vit.py:227-247)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.img_size = img_size
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def block_tensors(blk) -> dict:
    """Reference parameter names of one block -> engine vocabulary."""
    return {
        "ln1_g": blk.norm1.weight, "ln1_b": blk.norm1.bias,
        "qkv_w": blk.attn.qkv.weight, "qkv_b": blk.attn.qkv.bias,
        "proj_w": blk.attn.proj.weight, "proj_b": blk.attn.proj.bias,
        "ln2_g": blk.norm2.weight, "ln2_b": blk.norm2.bias,
        "fc1_w": blk.mlp.fc1.weight, "fc1_b": blk.mlp.fc1.bias,
        "fc2_w": blk.mlp.fc2.weight, "fc2_b": blk.mlp.fc2.bias,
    }


def resolve_keep_rates(keep_rate_list, blocks):
    """keep_rate_list[idx] or the block default (reference models_vit.py:366,101-102)."""
    rates = []
    for i, blk in enumerate(blocks):
        kr = keep_rate_list[i] if keep_rate_list is not None else None
        rates.append(blk.attn.default_keep_rate if kr is None else float(kr))
    return rates


class VisionTransformer(nn.Module):
    """Vision Transformer with global average pooling and TopK token pruning
    (reference models_vit.py:253-527), computed by libtpat.so."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, hybrid_backbone=None, norm_layer=nn.LayerNorm, mean_pooling=False, mask_2d=True,
                 target_length=None, drop_loc: tuple = None, base_keep_rate: tuple = None,
                 precision: Optional[str] = None, fuse_token: bool = False, **kwargs):
        super().__init__()
        assert hybrid_backbone is None, "hybrid backbones are not part of the pruning path"
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.num_extra_tokens = 1
        self.num_heads = num_heads
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        keep_rate_list = [1.0] * depth
        for drop_loc_idx in (drop_loc or ()):
            keep_rate_list[drop_loc_idx] = base_keep_rate                   # models_vit.py:283-285
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, block_id=i,
                  default_keep_rate=keep_rate_list[i])
            for i in range(depth)])
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)

        assert mean_pooling == True                                          # models_vit.py:307
        self.mean_pooling = mean_pooling
        self.fc_norm = norm_layer(embed_dim)
        self.mask_2d = mask_2d
        self.target_length = target_length
        self.use_custom_rank = None
        self.retain_max = None
        self.retain_min = None
        self.drop_token_blk_idx = None

        # EViT fused inattentive token (NOT in the reference forward; parity unpinned): a block that drops
        # tokens appends sum(score * dropped tokens) as one extra token after the kept ones.
        self.fuse_token = bool(fuse_token)
        self.precision = resolve_precision(precision)
        self.use_cuda_graph = False
        self.graph_static_io = False     # with use_cuda_graph: replay on the caller's input buffer, outputs as views (engine._run_graph)
        self._mlp_hidden = int(embed_dim * mlp_ratio)
        self._engines = EnginePool(_lib.VARIANT_AUDIOMAE, depth, embed_dim, num_heads, self._mlp_hidden)
        self.last_scores = None      # device tensors of the most recent forward
        self.last_topk_idx = None

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # ---- engine plumbing -----------------------------------------------------------------
    def _engine_tensors(self):
        return {
            "patch_w": self.patch_embed.proj.weight, "patch_b": self.patch_embed.proj.bias,
            "extra_tok": self.cls_token, "pos": self.pos_embed,
            "blocks": [block_tensors(b) for b in self.blocks],
            "norm_g": self.fc_norm.weight, "norm_b": self.fc_norm.bias,
            "head_ln_g": None, "head_ln_b": None,
            "head_w": self.head.weight, "head_b": self.head.bias,
        }

    def _pack_key(self):
        return param_key(self)

    def _device_of_params(self):
        return self.cls_token.device

    @property
    def _engine(self):
        """The ForwardEngine of the device this (replica of the) model lives on."""
        return self._engines.get(self._device_of_params())

    def invalidate_packed(self):
        """Call after changing weights through ``p.data`` (no version bump): drops bf16 copies and captured graphs."""
        self._engines.invalidate()

    def _check_supported(self, x):
        if self.embed_dim != 64 * self.num_heads or tuple(self.patch_embed.patch_size) != (16, 16):
            raise NotImplementedError("libtpat supports 16x16 patches and head_dim 64 (ViT-S/B/L)")
        if tuple(self.patch_embed.proj.weight.shape[1:]) != (1, 16, 16):
            raise NotImplementedError("libtpat expects the 1-channel 16x16 patch projection "
                                      "(replace model.patch_embed as main_finetune.py:378 does)")
        if not isinstance(self.head, nn.Linear):
            raise NotImplementedError("num_classes == 0 (Identity head) is not supported")

    def random_masking_2d_indices(self, B, device, mask_t_prob, mask_f_prob, noise=None):
        """Patch-token indices [B, T'*F'] kept by ``random_masking_2d`` (models_vit.py:425-463): per clip, the
        ``int(T*(1-mask_t_prob))`` time columns and ``int(F*(1-mask_f_prob))`` frequency rows with the smallest
        noise, in argsort order; kept token (t', f') is original token ids_t[t'] * F + ids_f[f'].  The two
        ``torch.rand`` calls are made in the reference's order on the input's device, so a seeded generator gives
        the reference's masks.  ``noise`` = (noise_t [B,T], noise_f [B,F]) overrides the draw (tests)."""
        T, F = self.target_length // 16, 8                                    # models_vit.py:436-437
        len_keep_T, len_keep_F = int(T * (1 - mask_t_prob)), int(F * (1 - mask_f_prob))
        noise_t = torch.rand(B, T, device=device) if noise is None else noise[0].to(device)
        ids_t = torch.argsort(noise_t, dim=1)[:, :len_keep_T]
        noise_f = torch.rand(B, F, device=device) if noise is None else noise[1].to(device)
        ids_f = torch.argsort(noise_f, dim=1)[:, :len_keep_F]
        return (ids_t[:, :, None] * F + ids_f[:, None, :]).reshape(B, len_keep_T * len_keep_F).contiguous()

    # ---- fine-tune step (training mode, autograd enabled) ---------------------------------------
    def _train_roles(self):
        top = {"patch_w": self.patch_embed.proj.weight, "patch_b": self.patch_embed.proj.bias,
               "extra_tok": self.cls_token, "extra_tok_first": self.cls_token, "pos": self.pos_embed,
               "norm_g": self.fc_norm.weight, "norm_b": self.fc_norm.bias, "head_ln_g": None, "head_ln_b": None,
               "head_w": self.head.weight, "head_b": self.head.bias}
        return top, [block_tensors(b) for b in self.blocks]

    def _train_entries(self, roles):
        """(stage, role, parameter) in backward order: head / final norm, blocks last to first, patch embedding."""
        top, blocks = roles
        depth = len(blocks)
        ent = [(depth + 1, r, top[r]) for r in ("head_w", "head_b", "norm_g", "norm_b")]
        for i in reversed(range(depth)):
            ent += [(i + 1, r, blocks[i][r]) for r in _lib.BLOCK_GRAD_NAMES]
        ent += [(0, "patch_w", top["patch_w"]), (0, "patch_b", top["patch_b"]), (0, "cls", self.cls_token), (0, "pos", self.pos_embed)]
        return ent

    def _forward_train(self, x, rates, mask_t_prob, mask_f_prob):
        """models_vit.py:502-522 in training mode: 2-D token masking (:425-497) while mask probabilities are set,
        DropPath (:149,198,205) with the block's stochastic-depth rate, everything recorded for the native backward."""
        from .train import drop_path_scales, run_train_step_forward
        B, _, T, F = x.shape
        engine = self._engines.get_train(self._device_of_params())
        roles = self._train_roles()
        engine.attach(self._train_entries(roles), {})
        keep_idx = None
        if mask_t_prob > 0.0 or mask_f_prob > 0.0:
            keep_idx = self.random_masking_2d_indices(B, x.device, mask_t_prob, mask_f_prob,
                                                      noise=getattr(self, "_mask_noise_override", None))
        scales = getattr(self, "_drop_scales_override", None)       # tests inject the oracle's draws (CPU generator)
        if scales is None:
            scales = drop_path_scales([b.drop_path_rate for b in self.blocks], B, x.device)
        with torch.cuda.device(x.device):
            logits, scores, idxs = run_train_step_forward(engine, self.cls_token, x.reshape(B, T, F), rates, self.num_classes,
                                                          self.precision, roles, drop_scales=scales, mask_keep_idx=keep_idx)
        self.last_scores, self.last_topk_idx = scores, idxs
        return logits

    def forward_features(self, x, keep_rate_list=None, flag_extract_features: bool = False):
        """``fc_norm(x[:, 1:].mean(1))`` after the 12 blocks (models_vit.py:334-396): the classifier input, [B, D] fp32 --
        or ``(outcome, feature_dict)`` in extract mode.  Computed by the same native call as ``forward`` (the head GEMM
        also runs; its logits are discarded)."""
        out = self.forward(x, keep_rate_list, flag_extract_features=flag_extract_features)
        if out is None:
            return None
        if self.use_custom_rank is not None or self.drop_token_blk_idx is not None:
            raise NotImplementedError("forward_features on the ablation paths: use forward()")
        pooled = self._engine.last_pooled
        return (pooled, out[1]) if flag_extract_features else pooled

    def forward(self, x, keep_rate_list: Union[list, tuple, type(None)] = None, mask_t_prob=0.0, mask_f_prob=0.0,
                flag_extract_features: bool = False):
        if (keep_rate_list is not None) and (len(keep_rate_list) != len(self.blocks)):
            raise ValueError(f"keep_rate should be a list/tuple of length {len(self.blocks)}, got {keep_rate_list}")
        self._check_supported(x)
        B, _, T, F = x.shape
        masking = mask_t_prob > 0.0 or mask_f_prob > 0.0
        if masking:
            assert flag_extract_features == False                             # models_vit.py:510
        else:
            assert T >= F and F == 128                                        # models_vit.py:336
        n_patches = (T // 16) * (F // 16)
        if self.pos_embed.shape[1] != n_patches + 1:
            raise RuntimeError(f"pos_embed has {self.pos_embed.shape[1]} rows but the input has {n_patches} patches + cls")
        rates = resolve_keep_rates(keep_rate_list, self.blocks)
        if self.training and torch.is_grad_enabled() and self.use_custom_rank is None and self.drop_token_blk_idx is None:
            assert flag_extract_features == False, "extract mode is an eval-time path"
            return self._forward_train(x, rates, mask_t_prob, mask_f_prob)
        self._engine.pack(self._engine_tensors, self._pack_key())
        spec = x.reshape(B, T, F)
        if masking:
            # fine-tune 2-D token masking (forward only; no autograd through libtpat yet): models_vit.py:425-497,509-512
            keep_idx = self.random_masking_2d_indices(B, x.device, mask_t_prob, mask_f_prob)
            logits, info = self._engine.run_stepwise(spec, rates, self.num_classes, precision=self.precision,
                                                     mask_keep_idx=keep_idx)
            self.last_scores, self.last_topk_idx = None, info["topk_idx"]
            return logits
        if self.use_custom_rank is not None or self.drop_token_blk_idx is not None:
            # ablation paths (models_vit.py:343-355,371-385): kernel-by-kernel forward
            if self.use_custom_rank is not None:
                assert flag_extract_features == False                         # models_vit.py:344
            if flag_extract_features:
                raise NotImplementedError("extract mode together with drop_token_blk_idx is not supported")
            logits, info = self._engine.run_stepwise(spec, rates, self.num_classes, precision=self.precision,
                                                     use_custom_rank=self.use_custom_rank,
                                                     drop_token_blk_idx=self.drop_token_blk_idx,
                                                     retain_min=self.retain_min, retain_max=self.retain_max)
            self.last_scores, self.last_topk_idx = None, info["topk_idx"]
            return logits                                                     # None when no token is retained (:384-385)
        self._engine.graph_static_io = self.graph_static_io
        logits, scores, idxs = self._engine.run(spec, rates, self.num_classes, want_all_scores=flag_extract_features,
                                                precision=self.precision, use_graph=self.use_cuda_graph, fuse_token=self.fuse_token)
        self.last_scores, self.last_topk_idx = scores, idxs
        if flag_extract_features:
            feature_dict = {'mel': x.cpu()}                                   # models_vit.py:338-339
            for i in range(len(self.blocks)):
                if scores[i] is not None:
                    feature_dict[f'block-{i}.attn_score'] = scores[i].cpu()   # :129
                if idxs[i] is not None:
                    feature_dict[f'block-{i}.topk_idx'] = idxs[i].cpu()       # :132
            return logits, feature_dict
        return logits


def vit_small_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_base_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_large_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_huge_patch14(**kwargs):
    return VisionTransformer(patch_size=14, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
