"""Drop-in mirror of the reference's AST pruning model API (ast/src/models/ast_models.py).

Same constructor signature, attribute names (``.v.blocks[i].attn.default_keep_rate`` ...) and
state-dict keys (``v.cls_token, v.dist_token, v.pos_embed, v.patch_embed.proj.*, v.blocks.*,
v.norm.*, v.head*, v.head_dist*, mlp_head.0.*, mlp_head.1.*``) as the reference; the forward is
one ``tpat_forward`` call into libtpat.so (AST flavour: two extra tokens, frequency-major token
order, CLS-row importance score -- SURVEY.md F2/F11/F12/F13).

    model = ASTModel(label_dim=527, input_tdim=1024, imagenet_pretrain=False, audioset_pretrain=False,
                     drop_loc=(3, 6, 9), base_keep_rate=0.7)
    logits = model(x)                                     # x [B, T, 128] on a B200
    logits, feats = model(x, flag_extract_features=True)

Stated differences: the reference needs ``timm==0.4.5`` for the DeiT skeleton and, with
``imagenet_pretrain=True``, a network download of DeiT weights; neither exists here, so the
skeleton is built locally and ``imagenet_pretrain=True`` raises -- load a checkpoint with
``load_state_dict`` instead.  ``@autocast`` (ast_models.py:424) is replaced by the explicit
``precision`` attribute ("bf16" tensor-core kernels / "fp32" parity kernels).  Inference only in
this round; the custom_rank / drop_token_blk_idx ablation paths run kernel by kernel (ForwardEngine.run_stepwise).
"""
import os
from functools import partial
from typing import Optional, Union

import torch
import torch.nn as nn

from . import _lib
from .engine import EnginePool, param_key, resolve_precision
from .models_vit import Block, PatchEmbed, block_tensors, resolve_keep_rates, trunc_normal_


class _DeiTDistilledSkeleton(nn.Module):
    """The attribute / parameter layout timm 0.4.5 gives ``vit_deit_base_distilled_patch16_384``
    (only what ASTModel touches: ast_models.py:273-330,395-409)."""

    def __init__(self, embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.0):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size=384, patch_size=16, in_chans=3, embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 2, embed_dim))
        self.pos_drop = nn.Dropout(p=0.0)
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=4.0, qkv_bias=True, drop_path=dpr[i],
                  norm_layer=norm_layer, block_id=i) for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, 1000)        # unused by the AST forward; kept for checkpoint keys
        self.head_dist = nn.Linear(embed_dim, 1000)   # unused by the AST forward; kept for checkpoint keys
        trunc_normal_(self.cls_token, std=.02)
        trunc_normal_(self.dist_token, std=.02)
        trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


class ASTModel(nn.Module):
    """The AST model with TopK token pruning (reference ast_models.py:239-508), computed by libtpat.so.

    :param label_dim: number of classes (527 AudioSet, 50 ESC-50, 35 Speech Commands v2)
    :param fstride, tstride: must be 16 (reference asserts the same, ast_models.py:258)
    :param input_fdim, input_tdim: mel bins / time frames of the input spectrogram
    :param drop_loc, base_keep_rate: pruning blocks (0-indexed) and their default keep rate
    """

    def __init__(self, label_dim=527, fstride=16, tstride=16, input_fdim=128, input_tdim=1024, imagenet_pretrain=True,
                 audioset_pretrain=False, model_size='base384', verbose=True, depth=12,
                 audioset_pretrained_model_path: str = None, drop_path_rate=0.0, drop_loc: tuple = None,
                 base_keep_rate: tuple = None, precision: Optional[str] = None, fuse_token: bool = False):
        super().__init__()
        assert fstride == 16 and tstride == 16, 'Currently only support fstride=16 and tstride=16.'
        if verbose:
            print('---------------AST Model Summary---------------')
            print('ImageNet pretraining: {:s}, AudioSet pretraining: {:s}'.format(str(imagenet_pretrain), str(audioset_pretrain)))
        if model_size != 'base384':
            raise Exception('We only support base384.')
        f_dim, t_dim = self.get_shape(fstride, tstride, input_fdim, input_tdim)
        num_patches = f_dim * t_dim

        if audioset_pretrain == False:
            if imagenet_pretrain:
                raise RuntimeError("imagenet_pretrain=True needs timm's DeiT download, which is unavailable offline; "
                                   "build with imagenet_pretrain=False and load a checkpoint with load_state_dict")
            self.v = _DeiTDistilledSkeleton(drop_path_rate=drop_path_rate)
            self.original_num_patches = self.v.patch_embed.num_patches
            self.oringal_hw = int(self.original_num_patches ** 0.5)
            self.original_embedding_dim = self.v.pos_embed.shape[2]
            self.mlp_head = nn.Sequential(nn.LayerNorm(self.original_embedding_dim),
                                          nn.Linear(self.original_embedding_dim, label_dim))
            self.v.patch_embed.num_patches = num_patches
            if verbose:
                print('frequncey stride={:d}, time stride={:d}'.format(fstride, tstride))
                print('number of patches={:d}'.format(num_patches))
            # 1-channel projection (ast_models.py:301-305) and a fresh learnable pos-embed (:326-330)
            self.v.patch_embed.proj = torch.nn.Conv2d(1, self.original_embedding_dim, kernel_size=(16, 16),
                                                      stride=(fstride, tstride))
            self.v.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 2, self.original_embedding_dim))
            trunc_normal_(self.v.pos_embed, std=.02)
        else:
            if imagenet_pretrain == False:
                raise ValueError('currently model pretrained on only audioset is not supported, please set '
                                 'imagenet_pretrain = True to use audioset pretrained model.')
            if audioset_pretrained_model_path is None or not os.path.exists(audioset_pretrained_model_path):
                raise FileNotFoundError(f"audioset_pretrained_model_path={audioset_pretrained_model_path!r} not found")
            sd = torch.load(audioset_pretrained_model_path, map_location='cpu')
            sd = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in sd.items()}  # DataParallel prefix
            base = ASTModel(label_dim=527, fstride=16, tstride=16, input_fdim=128, input_tdim=1024,
                            imagenet_pretrain=False, audioset_pretrain=False, model_size=model_size, verbose=False,
                            drop_loc=drop_loc, base_keep_rate=base_keep_rate, precision=precision, fuse_token=fuse_token)
            base.load_state_dict(sd, strict=True)
            self.v = base.v
            self.original_embedding_dim = self.v.pos_embed.shape[2]
            self.mlp_head = nn.Sequential(nn.LayerNorm(self.original_embedding_dim),
                                          nn.Linear(self.original_embedding_dim, label_dim))
            self.v.patch_embed.num_patches = num_patches
            if verbose:
                print('frequncey stride={:d}, time stride={:d}'.format(fstride, tstride))
                print('number of patches={:d}'.format(num_patches))
            # crop the AudioSet (8 x 64) positional embedding in time (ast_models.py:369-388)
            new_pos_embed = self.v.pos_embed[:, 2:, :].detach().reshape(1, 512, 768).transpose(1, 2).reshape(1, 768, 8, 64)
            if t_dim < 64:
                new_pos_embed = new_pos_embed[:, :, :, 32 - int(t_dim / 2): 32 - int(t_dim / 2) + t_dim]
            elif t_dim > 64:
                raise ValueError(f'{t_dim=} > 64')
            assert f_dim == 8
            new_pos_embed = new_pos_embed.reshape(1, 768, num_patches).transpose(1, 2)
            self.v.pos_embed = nn.Parameter(torch.cat([self.v.pos_embed[:, :2, :].detach(), new_pos_embed], dim=1))

        # TopK wiring (ast_models.py:391-412)
        self.use_custom_rank = None
        self.drop_token_blk_idx = None
        self.retain_min = None
        self.retain_max = None
        assert depth == len(self.v.blocks), "the DeiT-base skeleton has 12 blocks"
        self.depth = depth
        self.num_extra_tokens = 2
        self.num_heads = 12
        keep_rate_list = [1.0] * depth
        for drop_loc_idx in (drop_loc or ()):
            keep_rate_list[drop_loc_idx] = base_keep_rate
        for blk_id in range(depth):
            self.v.blocks[blk_id].attn.num_extra_tokens = 2
            self.v.blocks[blk_id].attn.block_id = blk_id
            self.v.blocks[blk_id].attn.default_keep_rate = keep_rate_list[blk_id]
            self.v.blocks[blk_id].block_id = blk_id
            self.v.blocks[blk_id].num_extra_tokens = 2

        self.label_dim = label_dim
        # EViT fused inattentive token (NOT in the reference forward; parity unpinned): a block that drops
        # tokens appends sum(score * dropped tokens) as one extra token after the kept ones.
        self.fuse_token = bool(fuse_token)
        self.precision = resolve_precision(precision)
        self.use_cuda_graph = False
        self.graph_static_io = False     # with use_cuda_graph: replay on the caller's input buffer, outputs as views (engine._run_graph)
        self._engines = EnginePool(_lib.VARIANT_AST, depth, 768, 12, 3072)
        self.last_scores = None
        self.last_topk_idx = None

    def get_shape(self, fstride, tstride, input_fdim=128, input_tdim=1024):
        """Output grid of the 16x16 conv (ast_models.py:416-422), computed arithmetically."""
        f_dim = (input_fdim - 16) // fstride + 1
        t_dim = (input_tdim - 16) // tstride + 1
        return f_dim, t_dim

    def _engine_tensors(self):
        v = self.v
        return {
            "patch_w": v.patch_embed.proj.weight, "patch_b": v.patch_embed.proj.bias,
            "extra_tok": torch.cat([v.cls_token.detach(), v.dist_token.detach()], dim=1),
            "pos": v.pos_embed,
            "blocks": [block_tensors(b) for b in v.blocks[:self.depth]],
            "norm_g": v.norm.weight, "norm_b": v.norm.bias,
            "head_ln_g": self.mlp_head[0].weight, "head_ln_b": self.mlp_head[0].bias,
            "head_w": self.mlp_head[1].weight, "head_b": self.mlp_head[1].bias,
        }

    def _pack_key(self):
        return param_key(self)

    def _device_of_params(self):
        return self.v.cls_token.device

    @property
    def _engine(self):
        """The ForwardEngine of the device this (replica of the) model lives on."""
        return self._engines.get(self._device_of_params())

    def invalidate_packed(self):
        """Call after changing weights through ``p.data`` (no version bump): drops bf16 copies and captured graphs."""
        self._engines.invalidate()

    # ---- fine-tune step (training mode, autograd enabled) ---------------------------------------
    def _train_roles(self):
        v = self.v
        top = {"patch_w": v.patch_embed.proj.weight, "patch_b": v.patch_embed.proj.bias,
               # cls and dist sit next to each other in the flat parameter buffer: one [2, D] extra-token table
               "extra_tok": v.cls_token, "extra_tok_first": v.cls_token, "pos": v.pos_embed,
               "norm_g": v.norm.weight, "norm_b": v.norm.bias,
               "head_ln_g": self.mlp_head[0].weight, "head_ln_b": self.mlp_head[0].bias,
               "head_w": self.mlp_head[1].weight, "head_b": self.mlp_head[1].bias}
        return top, [block_tensors(b) for b in v.blocks[:self.depth]]

    def _train_entries(self, roles):
        top, blocks = roles
        depth = len(blocks)
        ent = [(depth + 1, r, top[r]) for r in ("head_w", "head_b", "head_ln_g", "head_ln_b", "norm_g", "norm_b")]
        for i in reversed(range(depth)):
            ent += [(i + 1, r, blocks[i][r]) for r in _lib.BLOCK_GRAD_NAMES]
        ent += [(0, "patch_w", top["patch_w"]), (0, "patch_b", top["patch_b"]), (0, "cls", self.v.cls_token),
                (0, "dist", self.v.dist_token), (0, "pos", self.v.pos_embed)]
        return ent

    def _forward_train(self, x, rates):
        """ast_models.py:424-508 in training mode (DropPath of timm 0.4.5), recorded for the native backward.  The
        unused DeiT heads (v.head, v.head_dist) get no gradient, as in the reference."""
        from .train import drop_path_scales, run_train_step_forward
        B = x.shape[0]
        if not (self.v.cls_token.requires_grad and self.v.dist_token.requires_grad):
            raise NotImplementedError("cls_token and dist_token must both be trainable (they share one extra-token table)")
        engine = self._engines.get_train(self._device_of_params())
        roles = self._train_roles()
        engine.attach(self._train_entries(roles), {})
        scales = getattr(self, "_drop_scales_override", None)
        if scales is None:
            scales = drop_path_scales([b.drop_path_rate for b in self.v.blocks[:self.depth]], B, x.device, timm_04_style=True)
        with torch.cuda.device(x.device):
            logits, scores, idxs = run_train_step_forward(engine, self.v.cls_token, x, rates, self.label_dim, self.precision,
                                                          roles, drop_scales=scales)
        self.last_scores, self.last_topk_idx = scores, idxs
        return logits

    def forward(self, x, keep_rate_list: Union[list, tuple, type(None)] = None, flag_extract_features: bool = False):
        """x [B, time_frame_num, frequency_bins], e.g. (12, 1024, 128) (ast_models.py:431)."""
        if (keep_rate_list is not None) and (len(keep_rate_list) != len(self.v.blocks)):
            raise ValueError(f"keep_rate should be a list/tuple of length {len(self.v.blocks)}, got {keep_rate_list}")
        B, T, F = x.shape
        n_patches = (T // 16) * (F // 16)
        if self.v.pos_embed.shape[1] != n_patches + 2:
            raise RuntimeError(f"pos_embed has {self.v.pos_embed.shape[1]} rows but the input has {n_patches} patches + 2")
        rates = resolve_keep_rates(keep_rate_list, self.v.blocks)
        if self.training and torch.is_grad_enabled() and self.use_custom_rank is None and self.drop_token_blk_idx is None:
            assert flag_extract_features == False, "extract mode is an eval-time path"
            return self._forward_train(x, rates)
        self._engine.pack(self._engine_tensors, self._pack_key())
        if self.use_custom_rank is not None or self.drop_token_blk_idx is not None:
            # ablation paths (ast_models.py:445-457,480-497): kernel-by-kernel forward
            if self.use_custom_rank is not None:
                assert flag_extract_features == False                         # ast_models.py:446
            if flag_extract_features:
                raise NotImplementedError("extract mode together with drop_token_blk_idx is not supported")
            logits, info = self._engine.run_stepwise(x, rates, self.label_dim, precision=self.precision,
                                                     use_custom_rank=self.use_custom_rank,
                                                     drop_token_blk_idx=self.drop_token_blk_idx,
                                                     retain_min=self.retain_min, retain_max=self.retain_max)
            self.last_scores, self.last_topk_idx = None, info["topk_idx"]
            return logits                                                     # None when no token is retained (:495-497)
        self._engine.graph_static_io = self.graph_static_io
        logits, scores, idxs = self._engine.run(x, rates, self.label_dim, want_all_scores=flag_extract_features,
                                                precision=self.precision, use_graph=self.use_cuda_graph, fuse_token=self.fuse_token)
        self.last_scores, self.last_topk_idx = scores, idxs
        if flag_extract_features:
            feature_dict = {'mel': x.unsqueeze(1).transpose(2, 3).cpu()}      # ast_models.py:434-439
            for i in range(self.depth):
                if scores[i] is not None:
                    feature_dict[f'block-{i}.attn_score'] = scores[i].cpu()
                if idxs[i] is not None:
                    feature_dict[f'block-{i}.topk_idx'] = idxs[i].cpu()
            return logits, feature_dict
        return logits
