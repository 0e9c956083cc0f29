"""TEST INFRASTRUCTURE ONLY.  Import the UNMODIFIED reference models from /root/reference.

Only usable in the build container: the GPU box has no /root/reference, so nothing that runs
under ``-m gpu``, ``smoke()`` or ``bench.py`` may call into this module.  It exists to (1) pin
``vit_oracle.py`` against the real reference and (2) generate ``tests/golden``.

AudioMAE (`audiomae/models_vit.py`) imports four timm names (models_vit.py:20,23); they are
provided by an in-memory shim.  AST (`ast/src/models/ast_models.py`) does not parse as shipped
(SyntaxError at line 140) and requires timm==0.4.5 plus wget; its source text is
read, the one-token fix is applied to the in-memory copy, and it is exec'd behind a shim that
provides a DeiT-distilled skeleton through ``timm.create_model``.  No reference file is copied
into this repository.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("TPAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "audiomae", "models_vit.py"))


class _DropPath(nn.Module):
    """Stochastic depth; identity in eval mode (the only mode the oracle uses)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _install_timm_shim():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_tpat_shim", False):
        return sys.modules["timm"]
    timm = types.ModuleType("timm")
    timm._tpat_shim = True
    timm.__version__ = "0.4.5"
    data = types.ModuleType("timm.data")
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = _DropPath
    layers.to_2tuple = _to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.Attention = None
    vt.Block = None
    vt.PatchEmbed = None
    timm.data, timm.models = data, models
    models.layers, models.vision_transformer = layers, vt

    def create_model(name, pretrained=False, **kw):
        # DeiT-base distilled 384 skeleton as timm 0.4.5 builds it; every block class is the
        # reference's own patched class (ast_models.py:264-268).
        assert name == "vit_deit_base_distilled_patch16_384" and not pretrained
        from functools import partial
        m = nn.Module()
        dim, depth, heads = 768, 12, 12
        m.patch_embed = vt.PatchEmbed(img_size=384, patch_size=16, in_chans=3, embed_dim=dim)
        m.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        m.dist_token = nn.Parameter(torch.zeros(1, 1, dim))
        m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 2, dim))
        m.pos_drop = nn.Dropout(p=0.0)
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        m.blocks = nn.ModuleList([
            vt.Block(dim=dim, num_heads=heads, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                     attn_drop=0.0, drop_path=0.0, norm_layer=norm_layer)
            for _ in range(depth)])
        m.norm = norm_layer(dim)
        torch.nn.init.trunc_normal_(m.cls_token, std=.02)
        torch.nn.init.trunc_normal_(m.dist_token, std=.02)
        torch.nn.init.trunc_normal_(m.pos_embed, std=.02)
        return m

    timm.create_model = create_model
    for name, mod in (("timm", timm), ("timm.data", data), ("timm.models", models),
                      ("timm.models.layers", layers), ("timm.models.vision_transformer", vt)):
        sys.modules[name] = mod
    if "wget" not in sys.modules:
        sys.modules["wget"] = types.ModuleType("wget")
    return timm


_CACHE = {}


def load_audiomae_module():
    """Return the reference ``models_vit`` module (unmodified source, imported in place)."""
    if "audiomae" not in _CACHE:
        assert reference_available(), f"{REFERENCE_ROOT} not present"
        _install_timm_shim()
        import importlib.util
        path = os.path.join(REFERENCE_ROOT, "audiomae", "models_vit.py")
        spec = importlib.util.spec_from_file_location("_ref_models_vit", path)
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
        _CACHE["audiomae"] = mod
    return _CACHE["audiomae"]


def load_ast_module():
    """Return the reference ``ast_models`` module with the line-140 one-token fix (F9) applied
    to an in-memory copy of its source."""
    if "ast" not in _CACHE:
        assert reference_available(), f"{REFERENCE_ROOT} not present"
        _install_timm_shim()
        path = os.path.join(REFERENCE_ROOT, "ast", "src", "models", "ast_models.py")
        with open(path, "r") as f:
            src = f.read()
        broken = "attn_score': attn_score = attn["
        assert src.count(broken) == 1, "reference AST source changed; re-check F9"
        src = src.replace(broken, "attn_score': attn[")
        mod = types.ModuleType("_ref_ast_models")
        mod.__file__ = path
        exec(compile(src, path, "exec"), mod.__dict__)
        _CACHE["ast"] = mod
    return _CACHE["ast"]


def build_audiomae(num_classes=527, target_length=1024, drop_loc=(3, 6, 9), base_keep_rate=0.7):
    """Reference model exactly as main_finetune.py:358-382 builds it (eval mode)."""
    mv = load_audiomae_module()
    m = mv.vit_base_patch16(num_classes=num_classes, drop_path_rate=0.1, mean_pooling=True, mask_2d=True,
                            target_length=target_length, drop_loc=tuple(drop_loc),
                            base_keep_rate=base_keep_rate)
    m.patch_embed = mv.PatchEmbed((target_length, 128), 16, 1, 768)
    n_patches = m.patch_embed.num_patches
    m.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 1, 768), requires_grad=False)
    return m.eval()


def build_ast(label_dim=527, input_tdim=1024, drop_loc=(3, 6, 9), base_keep_rate=0.7):
    """Reference ASTModel as run.py:197-201 builds it without pretrained weights (eval mode)."""
    am = load_ast_module()
    m = am.ASTModel(label_dim=label_dim, fstride=16, tstride=16, input_fdim=128, input_tdim=input_tdim,
                    imagenet_pretrain=False, audioset_pretrain=False, model_size="base384", verbose=False,
                    drop_loc=tuple(drop_loc), base_keep_rate=base_keep_rate)
    return m.eval()
