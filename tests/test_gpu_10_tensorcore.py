"""GPU parity of the tcgen05 kernels (GEMM, fused attention) against (a) a PyTorch fp64 reference
computed from the SAME bf16-rounded operands and (b) the CUDA-core kernels of this library.
Tolerance: fp32 accumulation of bf16 products -> 2e-5 relative to the largest output before the
output is rounded; bf16 outputs add one bf16 rounding (2^-8 relative)."""
import pytest
import torch
import torch.nn.functional as F

import conftest  # noqa: F401
from gpu_util import dev, rel_err, ref_attention, ref_score

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K)
    (128, 256, 64), (300, 768, 768), (1026, 2304, 768), (4104, 3072, 768), (2890, 768, 3072), (1024, 768, 256),
    (25, 2304, 768), (32832, 768, 768), (257, 384, 384),
]


def _mk(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(dev()).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev()).to(torch.bfloat16)
    bias = (torch.randn(N, generator=g) * 0.1).to(dev())
    return a, w, bias


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tc_bias_and_gelu(M, N, K):
    from tpat import ops, _lib
    a, w, bias = _mk(M, N, K, 11)
    base = a.double() @ w.double().T + bias.double()
    out = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_TC)
    assert rel_err(out, base) < 2e-5
    out = ops.gemm(a, w, bias, torch.bfloat16, _lib.EPI_BIAS, _lib.IMPL_TC)
    assert rel_err(out.float(), base) < 5e-3
    out = ops.gemm(a, w, bias, torch.bfloat16, _lib.EPI_BIAS_GELU, _lib.IMPL_TC)
    assert rel_err(out.float(), F.gelu(base)) < 5e-3
    # cross-check with this library's CUDA-core kernel on identical operands
    simt = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_SIMT)
    tc = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_TC)
    assert rel_err(tc, simt) < 2e-5


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (2890, 768, 3072), (32832, 768, 768)])
def test_gemm_tc_residual_in_place(M, N, K):
    from tpat import ops, _lib
    a, w, bias = _mk(M, N, K, 12)
    res = torch.randn(M, N, generator=torch.Generator().manual_seed(13)).to(dev())
    ref = res.double() + a.double() @ w.double().T + bias.double()
    x = res.clone()
    out = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x, out=x)
    assert out.data_ptr() == x.data_ptr()
    assert rel_err(x, ref) < 2e-5


@pytest.mark.parametrize("B,P,extra", [(5, 512, 1), (3, 64, 2), (3, 104, 1), (2, 8, 2)])
def test_gemm_tc_patch_pos_epilogue(B, P, extra):
    """P % 32 == 0: TMA read-modify-write epilogue (position block in, shifted C block out); otherwise the register path."""
    from tpat import ops, _lib
    D = 768
    a, w, bias = _mk(B * P, D, 256, 14)
    pos = torch.randn(extra + P, D, generator=torch.Generator().manual_seed(15)).to(dev())
    out = torch.full((B * (extra + P), D), 7.0, device=dev())
    ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_POS, _lib.IMPL_TC, out=out, pos=pos, P=P, num_extra=extra)
    ref = (a.double() @ w.double().T + bias.double()).reshape(B, P, D) + pos[extra:].double()
    o = out.reshape(B, extra + P, D)
    assert rel_err(o[:, extra:], ref) < 2e-5
    assert torch.all(o[:, :extra] == 7.0)


@pytest.mark.parametrize("M,K", [(1000, 768), (513 * 3, 3072), (64, 768)])
def test_gemm_ln_fold_producer_and_consumer(M, K):
    """LayerNorm fold (tpat_gemm_ln): the residual GEMM emits bf16(x) + per-chunk partial moments (TMA epilogue for
    K = 768, register epilogue for K = 3072); a following GEMM with gamma-scaled weights normalises in its epilogue.
    Compared with LayerNorm(x) @ W^T + b in float64 and with the unfused tpat path (LayerNorm kernel -> bf16 -> GEMM)."""
    from tpat import ops, _lib
    D, N2 = 768, 2304
    g = torch.Generator().manual_seed(M + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev()).to(torch.bfloat16)
    w = (torch.randn(D, K, generator=g) * 0.03).to(dev()).to(torch.bfloat16)
    bias = (torch.randn(D, generator=g) * 0.1).to(dev())
    x0 = (torch.randn(M, D, generator=g) * 2.0 + 0.3).to(dev())
    x_ref = x0.clone()
    ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_TC, residual=x_ref, out=x_ref)
    x = x0.clone()
    _, xb, part = ops.gemm_ln(a, w, bias, torch.float32, _lib.EPI_BIAS_RESIDUAL, residual=x, out=x, emit=True)
    assert torch.equal(x, x_ref)                                   # C itself is unchanged by the extra outputs
    assert torch.equal(xb, x.to(torch.bfloat16))                   # bf16 copy = round-to-nearest of the fp32 value
    xc = x.double().reshape(M, D // 32, 32)
    assert torch.allclose(part[..., 0].double(), xc.sum(-1), rtol=1e-5, atol=1e-4)
    m2 = ((xc - xc.mean(-1, keepdim=True)) ** 2).sum(-1)
    assert torch.allclose(part[..., 1].double(), m2, rtol=1e-4, atol=1e-4)
    # consumer: y = LN(x) @ W2^T + b2, with gamma folded into W2 and beta into the bias
    gamma = (1.0 + 0.2 * torch.randn(D, generator=g)).to(dev())
    beta = (0.1 * torch.randn(D, generator=g)).to(dev())
    w2 = (torch.randn(N2, D, generator=g) * 0.03).to(dev())
    b2 = (torch.randn(N2, generator=g) * 0.1).to(dev())
    w2f = (w2 * gamma[None, :]).to(torch.bfloat16)
    colsum = w2f.float().sum(1).contiguous()
    b2f = (w2 @ beta + b2).contiguous()
    ref = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-6) @ w2.double().T + b2.double()
    for epi in (_lib.EPI_BIAS, _lib.EPI_BIAS_GELU):
        want = ref if epi == _lib.EPI_BIAS else torch.nn.functional.gelu(ref)
        got = ops.gemm_ln(xb, w2f, b2f, torch.bfloat16, epi, ln_part=part, ln_colsum=colsum, ln_eps=1e-6)
        y = ops.layernorm(x, gamma, beta, 1e-6, torch.bfloat16)
        unfused = ops.gemm(y, w2.to(torch.bfloat16), b2, torch.bfloat16, epi, _lib.IMPL_TC)
        e_fold, e_unf = rel_err(got.float(), want), rel_err(unfused.float(), want)
        print(f"[ln fold] M={M} K={K} epi={epi}: fold err {e_fold:.2e}, unfused err {e_unf:.2e}")
        assert e_fold < 1e-2 and e_fold < 1.5 * e_unf + 1e-3


def test_gemm_tc_rejects_bad_arguments():
    from tpat import ops, _lib
    a, w, bias = _mk(64, 64, 96, 16)
    with pytest.raises(RuntimeError, match="K %% 64|K % 64"):
        ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_TC)
    with pytest.raises(RuntimeError, match="bf16"):
        ops.gemm(a.float(), w.float(), bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_TC)


ATT_CASES = [(66, 2, "cls"), (25, 2, "cls"), (514, 2, "cls"), (361, 2, "cls"), (513, 1, "colmean"), (360, 1, "colmean"),
             (253, 1, "colmean"), (178, 1, "colmean"), (129, 1, "colmean"), (128, 1, "none"), (513, 1, "none")]


@pytest.mark.parametrize("N,extra,mode", ATT_CASES)
def test_attention_tc(N, extra, mode):
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(21)
    B, H = 3, 12
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    smode = {"cls": _lib.SCORE_CLS_ROW, "colmean": _lib.SCORE_COLMEAN, "none": _lib.SCORE_NONE}[mode]
    out, partial = ops.attention(qkv, B, N, H, extra, smode, _lib.IMPL_TC)
    ref_out, attn = ref_attention(qkv, B, N, H, extra)
    # P is rounded to bf16 before P.V and the output is bf16: two bf16 roundings
    assert rel_err(out.float(), ref_out) < 1e-2
    simt_out, simt_partial = ops.attention(qkv, B, N, H, extra, smode, _lib.IMPL_SIMT)
    assert rel_err(out.float(), simt_out.float()) < 1e-2
    if mode != "none":
        div = H if mode == "cls" else H * (N - extra)
        score, _ = ops.score_topk(partial, div, extra, 0)
        ref = ref_score(attn, extra, "cls" if mode == "cls" else "colmean")
        # scores are accumulated from fp32 probabilities (ex2.approx): ~1e-6 relative
        assert rel_err(score, ref) < 2e-5
        score_simt, _ = ops.score_topk(simt_partial, div, extra, 0)
        assert rel_err(score, score_simt) < 2e-5


@pytest.mark.parametrize("N,extra", [(513, 1), (514, 2), (200, 1)])
def test_attention_tc_online_softmax_rescale(N, extra):
    """Single-pass tiles: keys whose scores grow with the key index force the lazily rescaled online
    softmax to raise its reference max (and rescale O in TMEM) several times per row."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(22)
    B, H = 2, 12
    x = torch.randn(B, N, 3, H, 64, generator=g) * 1.5
    ramp = 1.0 + 24.0 * torch.arange(N, dtype=torch.float32) / N
    x[:, :, 1] *= ramp[None, :, None, None]                      # K rows get larger with the key index
    qkv = x.reshape(B * N, 3 * H * 64).to(dev()).to(torch.bfloat16)
    out, _ = ops.attention(qkv, B, N, H, extra, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, attn = ref_attention(qkv, B, N, H, extra)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2
    # the same inputs through the two-pass (score) tiles must agree with the single-pass tiles
    out2, partial = ops.attention(qkv, B, N, H, extra, _lib.SCORE_COLMEAN, _lib.IMPL_TC)
    assert rel_err(out.float(), out2.float()) < 1e-2
    score, _ = ops.score_topk(partial, H * (N - extra), extra, 0)
    assert rel_err(score, ref_score(attn, extra, "colmean")) < 2e-5


@pytest.mark.parametrize("M,N,K", [(3100, 768, 256), (5001, 2304, 768), (4100, 800, 128), (32832, 3072, 768)])
@pytest.mark.parametrize("with_bias", [True, False])
def test_gemm_tma_store_epilogue_is_bit_identical(M, N, K, with_bias, monkeypatch):
    """bf16 outputs of the bias / bias + GELU epilogues of the CTA-pair kernel leave through TMA stores (32 x 32 blocks staged
    in 64B-swizzled shared memory).  Same bits as the per-lane-store epilogue (TPAT_GEMM_TMA_STORE=0), row tails clipped
    by the tensor map (M % 32 != 0), column tiles that end inside a 256-column tile (N = 800), nothing written past M."""
    from tpat import ops, _lib
    a, w, bias = _mk(M, N, K, 13)
    b = bias if with_bias else None
    base = a.double() @ w.double().T + (bias.double() if with_bias else 0.0)
    for epi, ref in ((_lib.EPI_BIAS, base), (_lib.EPI_BIAS_GELU, F.gelu(base))):
        big = torch.full((M + 64, N), 7.0, device=dev(), dtype=torch.bfloat16)
        monkeypatch.setenv("TPAT_GEMM_TMA_STORE", "1")
        out = ops.gemm(a, w, b, torch.bfloat16, epi, _lib.IMPL_TC, out=big[:M])
        assert rel_err(out.float(), ref) < 5e-3
        assert (big[M:] == 7.0).all()
        monkeypatch.setenv("TPAT_GEMM_TMA_STORE", "0")
        old = ops.gemm(a, w, b, torch.bfloat16, epi, _lib.IMPL_TC)
        assert torch.equal(out, old)
