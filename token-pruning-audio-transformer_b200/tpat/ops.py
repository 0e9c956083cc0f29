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


# ---- fine-tune step: per-kernel wrappers (tests / hand assembly; the model classes use train.TrainEngine) ----------

_partials_ws = {}


def _partials(device) -> torch.Tensor:
    t = _partials_ws.get(device)
    if t is None:
        t = _partials_ws[device] = torch.empty(lib.tpat_bwd_partials_floats(4096), device=device, dtype=torch.float32)
    return t


def gemm_f32(a: torch.Tensor, b: torch.Tensor, trans_a: int = 0, trans_b: int = 0, out: Optional[torch.Tensor] = None,
             accumulate: bool = False) -> torch.Tensor:
    """out (+)= op(a) @ op(b), fp32 CUDA cores.  trans_a: a is stored [K, M]; trans_b: b is stored [N, K]."""
    _req(a, torch.float32, "a"); _req(b, torch.float32, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    check(lib.tpat_gemm_f32(a.data_ptr(), a.shape[1], int(trans_a), b.data_ptr(), b.shape[1], int(trans_b), out.data_ptr(), N, M, N, K,
                            1 if accumulate else 0, _stream()), "tpat_gemm_f32")
    return out


def transpose(x: torch.Tensor, out_dtype: torch.dtype, ld_dst: Optional[int] = None) -> torch.Tensor:
    """x [rows, cols] -> [cols, ld_dst] (zero-padded beyond rows), cast to out_dtype."""
    _req(x, name="x")
    rows, cols = x.shape
    ld = rows if ld_dst is None else ld_dst
    out = torch.empty(cols, ld, device=x.device, dtype=out_dtype)
    check(lib.tpat_transpose(x.data_ptr(), _DT[x.dtype], cols, out.data_ptr(), _DT[out_dtype], ld, rows, cols, _stream()), "tpat_transpose")
    return out


def row_bwd(dy: Optional[torch.Tensor], x: Optional[torch.Tensor], gamma: Optional[torch.Tensor], g_up: Optional[torch.Tensor],
            op_dtype: torch.dtype, row_scale: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None, n_in: int = 0,
            extra: int = 0, eps: float = 1e-6, src_offset: int = 0, rows_out: Optional[int] = None):
    """tpat_row_bwd: (g_out fp32, gb op_dtype, dgamma, dbeta, dbias).  With ``idx`` [B, k] the rows are scattered back to
    [B, extra + n_in, D] (backward of the token gather)."""
    ref = dy if dy is not None else g_up
    B, rows_src, D = ref.shape
    inv = None
    if idx is not None:
        inv = torch.empty(B, n_in, device=ref.device, dtype=torch.int32)
        check(lib.tpat_inverse_index(idx.contiguous().data_ptr(), inv.data_ptr(), B, n_in, idx.shape[1], _stream()), "tpat_inverse_index")
        r_out = extra + n_in
    else:
        r_out = rows_out if rows_out is not None else rows_src - src_offset
    g_out = torch.empty(B, r_out, D, device=ref.device, dtype=torch.float32)
    gb = torch.empty(B, r_out, D, device=ref.device, dtype=op_dtype)
    dg = torch.zeros(D, device=ref.device); db = torch.zeros(D, device=ref.device); dbias = torch.zeros(D, device=ref.device)
    check(lib.tpat_row_bwd(_ptr(dy), _DT[dy.dtype] if dy is not None else _lib.F32, _ptr(x), _ptr(gamma), _ptr(g_up), g_out.data_ptr(),
                           gb.data_ptr(), _DT[op_dtype], _ptr(row_scale), _ptr(inv), _partials(ref.device).data_ptr(), dg.data_ptr(),
                           db.data_ptr(), dbias.data_ptr(), B, rows_src, r_out, extra, src_offset, D, float(eps), _stream()), "tpat_row_bwd")
    return g_out, gb, dg, db, dbias


def colsum(x: torch.Tensor) -> torch.Tensor:
    _req(x, name="x")
    M, C = x.shape
    out = torch.zeros(C, device=x.device, dtype=torch.float32)
    check(lib.tpat_colsum(x.data_ptr(), _DT[x.dtype], C, M, C, _partials(x.device).data_ptr(), out.data_ptr(), _stream()), "tpat_colsum")
    return out


def batch_sum(x: torch.Tensor) -> torch.Tensor:
    """x [B, n] fp32 -> sum over dim 0."""
    _req(x, torch.float32, "x")
    B, n = x.shape
    out = torch.empty(n, device=x.device, dtype=torch.float32)
    check(lib.tpat_batch_sum(x.data_ptr(), out.data_ptr(), B, n, n, 0, _stream()), "tpat_batch_sum")
    return out


def pool_norm_bwd(x: torch.Tensor, dpooled: torch.Tensor, variant: int, g1, b1, eps1: float, g2=None, eps2: float = 0.0):
    _req(x, torch.float32, "x"); _req(dpooled, torch.float32, "dpooled")
    B, N, D = x.shape
    dx = torch.empty_like(x)
    outs = [torch.zeros(D, device=x.device) for _ in range(4)]
    check(lib.tpat_pool_norm_bwd(x.data_ptr(), dpooled.data_ptr(), dx.data_ptr(), g1.data_ptr(), b1.data_ptr(), float(eps1), _ptr(g2),
                                 float(eps2), _partials(x.device).data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                 outs[3].data_ptr(), B, N, D, variant, _stream()), "tpat_pool_norm_bwd")
    return (dx, *outs)


def attention_train(qkv: torch.Tensor, B: int, N: int, H: int, num_extra: int, score_mode: int, impl: int):
    """Training forward of the fused attention: (out, lse [B, H, N] fp32)."""
    _req(qkv, name="qkv")
    out = torch.empty(B * N, H * 64, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(B, H, N, device=qkv.device, dtype=torch.float32)
    partial = None
    if score_mode == _lib.SCORE_CLS_ROW:
        partial = torch.empty(B, H, N, device=qkv.device, dtype=torch.float32)
    elif score_mode == _lib.SCORE_COLMEAN:
        partial = torch.empty(B, H * attention_qtiles(N, impl), N, device=qkv.device, dtype=torch.float32)
    check(lib.tpat_attention_train(qkv.data_ptr(), out.data_ptr(), _DT[qkv.dtype], _ptr(partial), score_mode, lse.data_ptr(), B, N, H,
                                   64, num_extra, 64 ** -0.5, impl, _stream()), "tpat_attention_train")
    return out, lse


def attention_bwd(qkv: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, lse: torch.Tensor, B: int, N: int, H: int, impl: int,
                  dbias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dqkv; ``dbias`` [3 * H * 64] fp32 (tcgen05 path only) additionally receives += the column sums of dqkv."""
    _req(qkv, name="qkv")
    if dbias is not None:
        _req(dbias, torch.float32, "dbias"); _req(out, qkv.dtype, "out"); _req(d_out, qkv.dtype, "d_out"); _req(lse, torch.float32, "lse")
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(lib.tpat_attention_bwd_ws_floats(B, N, H, 64), device=qkv.device, dtype=torch.float32)
    check(lib.tpat_attention_bwd(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), _DT[qkv.dtype], B, N,
                                 H, 64, 64 ** -0.5, impl, delta.data_ptr(), _ptr(dbias), _stream()), "tpat_attention_bwd")
    return dqkv


def gemm_train(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out_dtype: torch.dtype, epilogue: int, impl: int,
               residual: Optional[torch.Tensor] = None, want_dact: bool = False, aux: Optional[torch.Tensor] = None,
               row_scale: Optional[torch.Tensor] = None, rows_per_clip: int = 0, w_kn: bool = False,
               colsum_out: Optional[torch.Tensor] = None):
    """tpat_gemm_train: out = epilogue(a @ w.T + bias) with the training extras (GELU-derivative output, GELU-backward
    epilogue, DropPath row scale).  ``w_kn``: ``w`` is [K, N] and out = epilogue(a @ w) (data gradients on the forward
    weight, tcgen05 path only)."""
    import ctypes
    _req(a, name="a"); _req(w, a.dtype, "w")
    M, K = a.shape
    N = w.shape[1] if w_kn else w.shape[0]
    assert w.shape[0 if w_kn else 1] == K
    out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    ex = _lib.GemmExtra()
    pre = None
    if want_dact:
        pre = torch.empty(M, N, device=a.device, dtype=out_dtype)
        ex.dact_out, ex.ld_dact = pre.data_ptr(), N
    if aux is not None:
        _req(aux, out_dtype, "aux")
        ex.aux, ex.ld_aux = aux.data_ptr(), N
    if row_scale is not None:
        ex.row_scale, ex.rows_per_clip = row_scale.data_ptr(), rows_per_clip
    ex.w_kn = 1 if w_kn else 0
    if colsum_out is not None:           # += column sums of the output (DGELU epilogue: the producing Linear's bias gradient)
        _req(colsum_out, torch.float32, "colsum_out")
        nws = max(int(lib.tpat_gemm_colsum_ws_floats(M, N)), int(lib.tpat_bwd_partials_floats(N)))
        ws = torch.empty(nws, device=a.device, dtype=torch.float32)
        ex.colsum_out, ex.colsum_ws, ex.colsum_ws_floats = colsum_out.data_ptr(), ws.data_ptr(), nws
    check(lib.tpat_gemm_train(a.data_ptr(), _DT[a.dtype], K, w.data_ptr(), _DT[w.dtype], _ptr(bias), out.data_ptr(), _DT[out_dtype], N,
                              _ptr(residual), N if residual is not None else 0, M, N, K, epilogue, impl, ctypes.byref(ex), _stream()),
          "tpat_gemm_train")
    return (out, pre) if want_dact else out


def adamw(p, g, m, v, p_bf16, chunks, groups, lr, beta1, beta2, eps, step, grad_scale=1.0, step_dev=None):
    check(lib.tpat_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p_bf16), chunks.data_ptr(), chunks.shape[0],
                         groups.data_ptr(), float(lr), float(beta1), float(beta2), float(eps), int(step), _ptr(step_dev),
                         float(grad_scale), _stream()), "tpat_adamw")


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out [Mo, No] fp32 += dy[K, Mo].T @ x[K, No] (bf16 operands, tcgen05 split-K kernel)."""
    _req(dy, torch.bfloat16, "dy"); _req(x, torch.bfloat16, "x")
    K, Mo = dy.shape
    No = x.shape[1]
    if out is None:
        out = torch.zeros(Mo, No, device=dy.device, dtype=torch.float32)
    check(lib.tpat_gemm_wgrad(dy.data_ptr(), Mo, x.data_ptr(), No, out.data_ptr(), No, K, Mo, No, _stream()), "tpat_gemm_wgrad")
    return out
