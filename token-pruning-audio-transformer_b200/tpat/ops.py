"""Per-kernel Python wrappers over the C-ABI (torch tensors in, torch tensors out).

torch is used for device memory and the current stream only; every computation happens in
libtpat.so.  These wrappers exist so that each kernel can be parity-tested in isolation and so
that a caller can assemble a forward by hand; the model classes use ``engine.ForwardEngine``
(one native call per forward) instead.
"""
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: libtpat has no CPU path")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def patchify(spec: torch.Tensor, out_dtype: torch.dtype, order: int, tokens: Optional[torch.Tensor] = None,
             extra_tok: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """spec [B,T,F] fp32 -> patch matrix [B*P, 256]; optionally fills the extra rows of ``tokens``."""
    _req(spec, torch.float32, "spec")
    B, T, F = spec.shape
    P = (T // 16) * (F // 16)
    patches = torch.empty(B * P, 256, device=spec.device, dtype=out_dtype)
    D = tokens.shape[-1] if tokens is not None else 0
    extra = extra_tok.shape[0] if extra_tok is not None else 0
    check(lib.tpat_patchify(spec.data_ptr(), patches.data_ptr(), _DT[out_dtype], _ptr(tokens), _ptr(extra_tok), _ptr(pos),
                            B, T, F, D, extra, order, _stream()), "tpat_patchify")
    return patches


def patch_stats(spec: torch.Tensor, order: int, want_mean: bool = True, want_std: bool = False):
    """spec [B,T,F] fp32 -> (mean [B,P] or None, std [B,P] or None): per-patch statistics in token order."""
    _req(spec, torch.float32, "spec")
    B, T, F = spec.shape
    P = (T // 16) * (F // 16)
    mean = torch.empty(B, P, device=spec.device, dtype=torch.float32) if want_mean else None
    std = torch.empty(B, P, device=spec.device, dtype=torch.float32) if want_std else None
    check(lib.tpat_patch_stats(spec.data_ptr(), _ptr(mean), _ptr(std), B, T, F, order, _stream()), "tpat_patch_stats")
    return mean, std


def gather_rank(rank: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """rank [B,n] fp32, idx [B,k] int64 -> rank gathered along dim 1."""
    _req(rank, torch.float32, "rank"); _req(idx, torch.int64, "idx")
    B, n = rank.shape
    k = idx.shape[1]
    out = torch.empty(B, k, device=rank.device, dtype=torch.float32)
    check(lib.tpat_gather_rank(rank.data_ptr(), idx.data_ptr(), out.data_ptr(), B, n, k, _stream()), "tpat_gather_rank")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out_dtype: torch.dtype,
              split3: bool = False) -> torch.Tensor:
    """LayerNorm over the last dim.  ``split3``: bf16 output [..., 3 * D] = [hi | lo | hi] (split-bf16 triple)."""
    _req(x, torch.float32, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    D = x.shape[-1]
    rows = x.numel() // D
    if split3:
        y = torch.empty(tuple(x.shape[:-1]) + (3 * D,), device=x.device, dtype=torch.bfloat16)
        dt = _lib.BF16_SPLIT3
    else:
        y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        dt = _DT[out_dtype]
    check(lib.tpat_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), dt, rows, D,
                             float(eps), _stream()), "tpat_layernorm")
    return y


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, 2 * cols] = [hi | lo]."""
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    out = torch.empty(rows, 2 * cols, device=x.device, dtype=torch.bfloat16)
    check(lib.tpat_split_bf16(x.data_ptr(), out.data_ptr(), rows, cols, _stream()), "tpat_split_bf16")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out_dtype: torch.dtype, epilogue: int = 0,
         impl: int = _lib.IMPL_SIMT, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         pos: Optional[torch.Tensor] = None, P: int = 0, num_extra: int = 0, k_cols: Optional[int] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias).  a [M,K], w [N,K] (nn.Linear layout).  ``k_cols``: use only the first k_cols
    columns of ``a`` (row pitch stays a.shape[1])."""
    _req(a, name="a"); _req(w, a.dtype, "w")
    M, lda = a.shape
    K = lda if k_cols is None else k_cols
    N = w.shape[0]
    if out is None:
        rows = M if epilogue != _lib.EPI_BIAS_POS else (M // P) * (P + num_extra)
        out = torch.empty(rows, N, device=a.device, dtype=out_dtype)
    _req(out, out_dtype, "out")
    if residual is not None:
        _req(residual, torch.float32, "residual")
    check(lib.tpat_gemm(a.data_ptr(), _DT[a.dtype], lda, w.data_ptr(), _DT[w.dtype], _ptr(bias), out.data_ptr(),
                        _DT[out_dtype], N, _ptr(residual), N if residual is not None else 0, _ptr(pos), P, num_extra,
                        M, N, K, epilogue, impl, _stream()), "tpat_gemm")
    return out


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out_dtype: torch.dtype, epilogue: int,
            residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, emit: bool = False,
            ln_part: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None, ln_eps: float = 1e-6):
    """tpat_gemm_ln on the tcgen05 path.  ``emit`` (residual epilogue): also return (xb bf16 [M,N], part [M,N/32,2]).
    ``ln_part`` / ``ln_colsum`` (bias, bias+GELU epilogues): fold the LayerNorm of the rows of ``a`` into the epilogue."""
    import ctypes
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    fold = _lib.LnFold()
    xb = part = None
    if emit:
        xb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
        part = torch.empty(M, N // 32, 2, device=a.device, dtype=torch.float32)
        fold.xb, fold.ldxb, fold.part_out = xb.data_ptr(), N, part.data_ptr()
    if ln_part is not None:
        _req(ln_part, torch.float32, "ln_part"); _req(ln_colsum, torch.float32, "ln_colsum")
        fold.ln_part, fold.ln_colsum, fold.ln_eps = ln_part.data_ptr(), ln_colsum.data_ptr(), float(ln_eps)
    check(lib.tpat_gemm_ln(a.data_ptr(), _lib.BF16, K, w.data_ptr(), _lib.BF16, _ptr(bias), out.data_ptr(), _DT[out_dtype], N,
                           _ptr(residual), N if residual is not None else 0, None, 0, 0, M, N, K, epilogue, _lib.IMPL_TC,
                           ctypes.byref(fold), _stream()), "tpat_gemm_ln")
    return (out, xb, part) if emit else out


def attention_qtiles(N: int, impl: int) -> int:
    return lib.tpat_attention_qtiles(N, impl)


def attention(qkv: torch.Tensor, B: int, N: int, H: int, num_extra: int, score_mode: int, impl: int,
              qk_planes: Optional[torch.Tensor] = None):
    """qkv [B*N, 3*H*64] -> (out [B*N, H*64], score_partial or None).  ``qk_planes`` [B*N, 4*H*64] bf16 = split-bf16
    [q_hi k_hi | q_lo k_lo] for the score tiles (tpat_attention_split)."""
    _req(qkv, name="qkv")
    if qk_planes is not None:
        _req(qk_planes, torch.bfloat16, "qk_planes")
    out = torch.empty(B * N, H * 64, device=qkv.device, dtype=qkv.dtype)
    partial = None
    if score_mode == _lib.SCORE_CLS_ROW:
        partial = torch.empty(B, H, N, device=qkv.device, dtype=torch.float32)
    elif score_mode == _lib.SCORE_COLMEAN:
        partial = torch.empty(B, H * attention_qtiles(N, impl), N, device=qkv.device, dtype=torch.float32)
    check(lib.tpat_attention_split(qkv.data_ptr(), _ptr(qk_planes), out.data_ptr(), _DT[qkv.dtype], _ptr(partial), score_mode,
                                   B, N, H, 64, num_extra, 64 ** -0.5, impl, _stream()), "tpat_attention")
    return out, partial


def score_topk(partial: torch.Tensor, divisor: float, num_extra: int, k: int, want_rest: bool = False):
    """partial [B,R,N] -> (score [B,N-extra] fp32, topk_idx [B,k] int64 or None[, rest_idx [B,n-k] int32])."""
    _req(partial, torch.float32, "partial")
    B, R, N = partial.shape
    score = torch.empty(B, N - num_extra, device=partial.device, dtype=torch.float32)
    idx = torch.empty(B, k, device=partial.device, dtype=torch.int64) if k > 0 else None
    rest = torch.empty(B, N - num_extra - k, device=partial.device, dtype=torch.int32) if want_rest else None
    check(lib.tpat_score_topk(partial.data_ptr(), R, float(divisor), score.data_ptr(), _ptr(idx), _ptr(rest), B, N, num_extra,
                              k, _stream()), "tpat_score_topk")
    return (score, idx, rest) if want_rest else (score, idx)


def gather_layernorm(x: torch.Tensor, idx: torch.Tensor, num_extra: int, gamma: Optional[torch.Tensor],
                     beta: Optional[torch.Tensor], eps: float, out_dtype: torch.dtype, score: Optional[torch.Tensor] = None,
                     rest_idx: Optional[torch.Tensor] = None):
    """x [B,N,D] fp32, idx [B,k] -> (x_out [B,extra+k(+1),D] fp32, LayerNorm(x_out) or None).
    With ``score`` and ``rest_idx`` the EViT fused inattentive token is appended as the last row."""
    _req(x, torch.float32, "x"); _req(idx, torch.int64, "idx")
    B, N, D = x.shape
    k = idx.shape[1]
    fuse = rest_idx is not None
    rows = num_extra + k + (1 if fuse else 0)
    xo = torch.empty(B, rows, D, device=x.device, dtype=torch.float32)
    yo = torch.empty(B, rows, D, device=x.device, dtype=out_dtype) if gamma is not None else None
    check(lib.tpat_gather_layernorm(x.data_ptr(), idx.data_ptr(), xo.data_ptr(), _ptr(gamma), _ptr(beta), _ptr(yo),
                                    _DT[out_dtype], B, N, k, num_extra, rows, D, float(eps), _stream()), "tpat_gather_layernorm")
    if fuse:
        _req(score, torch.float32, "score"); _req(rest_idx, torch.int32, "rest_idx")
        check(lib.tpat_fuse_token(x.data_ptr(), score.data_ptr(), rest_idx.data_ptr(), xo.data_ptr(), _ptr(gamma), _ptr(beta),
                                  _ptr(yo), _DT[out_dtype], B, N, rest_idx.shape[1], rows, rows - 1, num_extra, D, float(eps),
                                  _stream()), "tpat_fuse_token")
    return xo, yo


def pool_norm(x: torch.Tensor, variant: int, g1, b1, eps1: float, g2=None, b2=None, eps2: float = 0.0) -> torch.Tensor:
    _req(x, torch.float32, "x")
    B, N, D = x.shape
    out = torch.empty(B, D, device=x.device, dtype=torch.float32)
    check(lib.tpat_pool_norm(x.data_ptr(), out.data_ptr(), g1.data_ptr(), b1.data_ptr(), float(eps1), _ptr(g2), _ptr(b2),
                             float(eps2), B, N, D, variant, _stream()), "tpat_pool_norm")
    return out


def head(pooled: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """logits = pooled @ w.T + bias (fp32)."""
    _req(pooled, torch.float32, "pooled"); _req(w, torch.float32, "w")
    B, D = pooled.shape
    C = w.shape[0]
    out = torch.empty(B, C, device=pooled.device, dtype=torch.float32)
    check(lib.tpat_head(pooled.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), B, D, C, _stream()), "tpat_head")
    return out
