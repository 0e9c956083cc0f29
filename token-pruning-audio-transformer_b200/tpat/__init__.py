"""tpat -- host side of the B200-native token-pruned audio ViT forward.

``models_vit`` mirrors the reference's AudioMAE API (audiomae/models_vit.py), ``ast_models`` the
AST API (ast/src/models/ast_models.py); both compute through libtpat.so (hand-written sm_100a
CUDA behind the C-ABI in include/tpat.h).  Importing this package loads the shared library and
fails loudly if it is missing: there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (loads libtpat.so, verifies every symbol of include/tpat.h)
from . import ops, engine, models_vit, ast_models, extract, frontend, schedule, train, optim, lr_decay  # noqa: F401
from .models_vit import VisionTransformer, vit_base_patch16  # noqa: F401
from .ast_models import ASTModel  # noqa: F401
from .frontend import FbankFrontend  # noqa: F401

__all__ = ["ops", "engine", "models_vit", "ast_models", "frontend", "VisionTransformer", "vit_base_patch16", "ASTModel",
           "FbankFrontend"]
