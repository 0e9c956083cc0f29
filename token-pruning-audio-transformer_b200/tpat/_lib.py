"""ctypes binding of libtpat.so (the C-ABI declared in include/tpat.h).

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be built the
import fails, and every compute call raises ``RuntimeError(tpat_last_error())`` on a non-zero
status.
"""
import ctypes
import os
import sys
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("TPAT_LIB_PATH") or os.path.join(_PKG_ROOT, "lib", "libtpat.so")   # override: kernel experiments

TPAT_MAX_DEPTH = 32
TPAT_VERSION = 11         # must equal TPAT_VERSION in include/tpat.h (checked at load)
F32, BF16, BF16_SPLIT3 = 0, 1, 2
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_BIAS_POS, EPI_DGELU = 0, 1, 2, 3, 4
IMPL_SIMT, IMPL_TC = 0, 1
SCORE_NONE, SCORE_CLS_ROW, SCORE_COLMEAN = 0, 1, 2
TOKENS_TIME_MAJOR, TOKENS_FREQ_MAJOR = 0, 1
VARIANT_AUDIOMAE, VARIANT_AST = 0, 1


class BlockWeights(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_g", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
        "qkv_w_ln", "qkv_colsum", "qkv_b_ln", "fc1_w_ln", "fc1_colsum", "fc1_b_ln", "qk_w_split")]


class LnFold(Structure):
    """tpat_ln_fold (include/tpat.h): producer outputs / consumer inputs of the folded LayerNorm."""
    _fields_ = [("xb", c_void_p), ("ldxb", c_int), ("part_out", c_void_p), ("ln_part", c_void_p), ("ln_colsum", c_void_p),
                ("ln_eps", c_float)]


class ForwardArgs(Structure):
    _fields_ = [
        ("variant", c_int), ("impl", c_int), ("B", c_int), ("T", c_int), ("F", c_int),
        ("depth", c_int), ("D", c_int), ("H", c_int), ("Dh", c_int), ("num_classes", c_int),
        ("prune", c_int * TPAT_MAX_DEPTH), ("keep", c_int * TPAT_MAX_DEPTH), ("fuse_token", c_int),
        ("want_all_scores", c_int), ("score32", c_int), ("ln_eps", c_float),
        ("patch_w", c_void_p), ("patch_b", c_void_p), ("extra_tok", c_void_p), ("pos", c_void_p),
        ("blocks", BlockWeights * TPAT_MAX_DEPTH),
        ("norm_g", c_void_p), ("norm_b", c_void_p), ("norm_eps", c_float),
        ("head_ln_g", c_void_p), ("head_ln_b", c_void_p), ("head_ln_eps", c_float),
        ("head_w", c_void_p), ("head_b", c_void_p),
        ("spec", c_void_p), ("logits", c_void_p),
        ("scores", c_void_p * TPAT_MAX_DEPTH), ("topk_idx", c_void_p * TPAT_MAX_DEPTH),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("pooled", c_void_p),
    ]


class GemmExtra(Structure):
    """tpat_gemm_extra (include/tpat.h)."""
    _fields_ = [("dact_out", c_void_p), ("ld_dact", c_int), ("aux", c_void_p), ("ld_aux", c_int), ("row_scale", c_void_p),
                ("rows_per_clip", c_int), ("w_kn", c_int), ("colsum_out", c_void_p), ("colsum_ws", c_void_p),
                ("colsum_ws_floats", c_size_t)]


BLOCK_GRAD_NAMES = ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_g", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")


class BlockGrads(Structure):
    _fields_ = [(n, c_void_p) for n in BLOCK_GRAD_NAMES]


class BlockWt(Structure):
    _fields_ = [(n, c_void_p) for n in ("qkv_wt", "proj_wt", "fc1_wt", "fc2_wt")]


class TrainArgs(Structure):
    """tpat_train_args (include/tpat.h)."""
    _fields_ = [
        ("fwd", ForwardArgs),
        ("saved", c_void_p), ("saved_bytes", c_size_t),
        ("drop_scale", (c_void_p * 2) * TPAT_MAX_DEPTH),
        ("mask_keep_idx", c_void_p), ("n_keep", c_int),
        ("dlogits", c_void_p),
        ("wt", BlockWt * TPAT_MAX_DEPTH),
        ("grads", BlockGrads * TPAT_MAX_DEPTH),
        ("d_patch_w", c_void_p), ("d_patch_b", c_void_p), ("d_extra_tok", c_void_p), ("d_pos", c_void_p),
        ("d_norm_g", c_void_p), ("d_norm_b", c_void_p), ("d_head_ln_g", c_void_p), ("d_head_ln_b", c_void_p),
        ("d_head_w", c_void_p), ("d_head_b", c_void_p),
        ("bwd_workspace", c_void_p), ("bwd_workspace_bytes", c_size_t),
    ]


# every symbol include/tpat.h declares: (restype, argtypes)
SIGNATURES = {
    "tpat_version": (c_int, []),
    "tpat_last_error": (c_char_p, []),
    "tpat_device_ok": (c_int, []),
    "tpat_patchify": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_void_p]),
    "tpat_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "tpat_gemm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                          c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_gemm_ln": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                             c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(LnFold), c_void_p]),
    "tpat_attention_qtiles": (c_int, [c_int, c_int]),
    "tpat_attention": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               c_int, c_void_p]),
    "tpat_split_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tpat_attention_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_float, c_int, c_void_p]),
    "tpat_score_topk": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_gather_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "tpat_fuse_token": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "tpat_pool_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_float, c_int,
                               c_int, c_int, c_int, c_void_p]),
    "tpat_head": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tpat_fbank": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                           c_void_p, c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_void_p]),
    "tpat_patch_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_gather_rank": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tpat_gemm_train": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                c_int, c_int, c_int, c_int, c_int, POINTER(GemmExtra), c_void_p]),
    "tpat_gemm_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_gemm_colsum_ws_floats": (c_size_t, [c_int, c_int]),
    "tpat_gemm_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_transpose": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_attention_train": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_float, c_int, c_void_p]),
    "tpat_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                   c_int, c_void_p, c_void_p, c_void_p]),
    "tpat_attention_bwd_ws_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tpat_bwd_partials_floats": (c_size_t, [c_int]),
    "tpat_inverse_index": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tpat_row_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "tpat_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tpat_batch_sum": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_int, c_int, c_void_p]),
    "tpat_pool_norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_float, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tpat_adamw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_float,
                           c_float, c_float, c_int, c_void_p, c_float, c_void_p]),
    "tpat_counter_inc": (c_int, [c_void_p, c_void_p]),
    "tpat_sizeof_train_args": (c_size_t, []),
    "tpat_train_saved_bytes": (c_size_t, [POINTER(TrainArgs)]),
    "tpat_train_bwd_workspace_bytes": (c_size_t, [POINTER(TrainArgs)]),
    "tpat_train_forward": (c_int, [POINTER(TrainArgs), c_void_p]),
    "tpat_train_backward": (c_int, [POINTER(TrainArgs), c_int, c_int, c_void_p]),
    "tpat_sizeof_forward_args": (c_size_t, []),
    "tpat_forward_workspace_bytes": (c_size_t, [POINTER(ForwardArgs)]),
    "tpat_forward": (c_int, [POINTER(ForwardArgs), c_void_p]),
    "tpat_forward_launch_count": (c_int, [POINTER(ForwardArgs)]),
}


def _load():
    # Rebuild-if-stale on every import (stamp-checked: a no-op when no source / header / flag changed), so that an
    # edited csrc/ or a pulled tree never runs against an old binary.  TPAT_LIB_PATH (kernel experiments) and a
    # machine without nvcc (the prebuilt .so travelled with the tree) skip it.
    have_src = os.path.isdir(os.path.join(_PKG_ROOT, "csrc"))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.environ.get("TPAT_LIB_PATH") and have_src and (os.path.exists(nvcc) or not os.path.exists(LIB_PATH)):
        sys.path.insert(0, _PKG_ROOT)
        try:
            import build as _build  # token-pruning-audio-transformer_b200/build.py
            _build.build(verbose=False)
        finally:
            sys.path.pop(0)
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing and could not be built; there is no fallback path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export what tpat.h declares
        fn.restype = res
        fn.argtypes = args
    if lib.tpat_version() != TPAT_VERSION:
        raise ImportError(f"libtpat.so is version {lib.tpat_version()}, tpat/_lib.py expects {TPAT_VERSION}")
    if lib.tpat_sizeof_forward_args() != ctypes.sizeof(ForwardArgs):
        raise ImportError("ForwardArgs layout does not match the tpat_forward_args compiled into libtpat.so")
    if lib.tpat_sizeof_train_args() != ctypes.sizeof(TrainArgs):
        raise ImportError("TrainArgs layout does not match the tpat_train_args compiled into libtpat.so")
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.tpat_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RuntimeError(f"libtpat {what} failed (status {status}): {last_error()}")
