"""GPU parity of the CUDA-core kernels (fp32 parity path and the HBM-bound kernels) against plain
PyTorch fp32/fp64 references of the same op.  All calls go through the C-ABI (tpat.ops -> ctypes)."""
import pytest
import torch
import torch.nn.functional as F

import conftest  # noqa: F401
from gpu_util import dev, rel_err, ref_attention, ref_score

pytestmark = pytest.mark.gpu


def _ops():
    from tpat import ops, _lib
    return ops, _lib


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("B,T", [(2, 128), (3, 1024), (1, 16)])
def test_patchify_matches_conv_unfold(order, B, T):
    ops, _lib = _ops()
    g = torch.Generator().manual_seed(1)
    spec = torch.randn(B, T, 128, generator=g).to(dev())
    D, extra = 768, 1 if order == 0 else 2
    P = (T // 16) * 8
    extra_tok = torch.randn(extra, D, generator=g).to(dev())
    pos = torch.randn(extra + P, D, generator=g).to(dev())
    tokens = torch.zeros(B, extra + P, D, device=dev())
    for dt in (torch.float32, torch.bfloat16):
        patches = ops.patchify(spec, dt, order, tokens, extra_tok, pos)
        img = spec[:, None] if order == 0 else spec[:, None].transpose(2, 3)   # [B,1,T,F] / [B,1,F,T]
        ref = F.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(B * P, 256)
        assert torch.equal(patches.float(), ref.to(dt).float())
    assert torch.equal(tokens[:, :extra], (extra_tok + pos[:extra])[None].expand(B, -1, -1))
    assert torch.count_nonzero(tokens[:, extra:]) == 0


@pytest.mark.parametrize("D", [768, 384, 1024])
@pytest.mark.parametrize("rows", [1, 77, 4104])
def test_layernorm(D, rows):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(rows, D, generator=g) * 3 + 1.5).to(dev())
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(dev())
    b = (0.05 * torch.randn(D, generator=g)).to(dev())
    ref = F.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-6)
    y = ops.layernorm(x, w, b, 1e-6, torch.float32)
    assert rel_err(y, ref) < 2e-6                      # fp32 tolerance
    yb = ops.layernorm(x, w, b, 1e-6, torch.bfloat16)
    assert torch.equal(yb, y.to(torch.bfloat16))       # bf16 output = rounded fp32 result


@pytest.mark.parametrize("M,N,K", [(130, 768, 768), (1026, 2304, 768), (257, 768, 3072), (64, 527, 768), (5, 35, 768)])
def test_gemm_simt_fp32_epilogues(M, N, K):
    ops, _lib = _ops()
    g = torch.Generator().manual_seed(3)
    a = torch.randn(M, K, generator=g).to(dev())
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev())
    bias = (torch.randn(N, generator=g) * 0.1).to(dev())
    res = torch.randn(M, N, generator=g).to(dev())
    base = a.double() @ w.double().T + bias.double()
    out = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS, _lib.IMPL_SIMT)
    assert rel_err(out, base) < 1e-5
    out = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_GELU, _lib.IMPL_SIMT)
    assert rel_err(out, F.gelu(base)) < 1e-5
    x = res.clone()
    out = ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_RESIDUAL, _lib.IMPL_SIMT, residual=x, out=x)   # in place
    assert rel_err(out, res.double() + base) < 1e-5
    out = ops.gemm(a, w, None, torch.bfloat16, _lib.EPI_BIAS, _lib.IMPL_SIMT)
    assert rel_err(out.float(), a.double() @ w.double().T) < 5e-3


def test_gemm_simt_patch_pos_epilogue():
    ops, _lib = _ops()
    g = torch.Generator().manual_seed(4)
    B, P, extra, D = 3, 64, 2, 768
    a = torch.randn(B * P, 256, generator=g).to(dev())
    w = (torch.randn(D, 256, generator=g) * 0.06).to(dev())
    bias = torch.randn(D, generator=g).to(dev())
    pos = torch.randn(extra + P, D, generator=g).to(dev())
    out = torch.full((B * (extra + P), D), 7.0, device=dev())
    ops.gemm(a, w, bias, torch.float32, _lib.EPI_BIAS_POS, _lib.IMPL_SIMT, out=out, pos=pos, P=P, num_extra=extra)
    ref = (a.double() @ w.double().T + bias.double()).reshape(B, P, D) + pos[extra:].double()
    o = out.reshape(B, extra + P, D)
    assert rel_err(o[:, extra:], ref) < 1e-5
    assert torch.all(o[:, :extra] == 7.0)              # extra rows are not touched by the GEMM


@pytest.mark.parametrize("N,extra,mode", [(66, 2, "cls"), (513, 1, "colmean"), (360, 1, "colmean"), (179, 2, "cls"),
                                          (25, 2, "cls"), (33, 1, "colmean")])
def test_attention_simt_fp32(N, extra, mode):
    ops, _lib = _ops()
    g = torch.Generator().manual_seed(5)
    B, H = 2, 12
    qkv = torch.randn(B * N, 3 * H * 64, generator=g).to(dev())
    smode = _lib.SCORE_CLS_ROW if mode == "cls" else _lib.SCORE_COLMEAN
    out, partial = ops.attention(qkv, B, N, H, extra, smode, _lib.IMPL_SIMT)
    ref_out, attn = ref_attention(qkv, B, N, H, extra)
    assert rel_err(out, ref_out) < 1e-5
    div = H if mode == "cls" else H * (N - extra)
    score, _ = ops.score_topk(partial, div, extra, 0)
    assert rel_err(score, ref_score(attn, extra, mode)) < 1e-5
    out2, none = ops.attention(qkv, B, N, H, extra, _lib.SCORE_NONE, _lib.IMPL_SIMT)
    assert none is None and torch.equal(out, out2)


def test_score_topk_order_ties_and_nan():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(6)
    B, R, N, extra = 4, 12, 514, 2
    partial = torch.rand(B, R, N, generator=g).to(dev())
    k = 359
    score, idx = ops.score_topk(partial, float(R), extra, k)
    ref_score_ = partial[:, :, extra:].double().sum(1) / R
    assert rel_err(score, ref_score_) < 1e-6
    # descending order of the kernel's own fp32 score, identical selection to torch.topk on it
    tv, ti = torch.topk(score, k, dim=1, largest=True, sorted=True)
    assert torch.equal(torch.gather(score, 1, idx), tv)
    assert idx.dtype == torch.int64 and idx.shape == (B, k)
    for a, b in zip(idx.tolist(), ti.tolist()):
        assert set(a) == set(b)
    # ties: lower index first; NaN ranks highest (torch.topk semantics)
    p = torch.zeros(1, 1, 10, device=dev())
    p[0, 0, :] = torch.tensor([9., 9., 1., 5., 5., 5., 0., 2., float("nan"), 5.])
    _, idx = ops.score_topk(p, 1.0, 2, 6)
    assert idx.tolist() == [[6, 1, 2, 3, 7, 5]]        # candidates are tokens 2..9 -> local indices 0..7


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gather_layernorm(out_dtype):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(7)
    B, N, D, extra, k = 3, 66, 768, 2, 45
    x = torch.randn(B, N, D, generator=g).to(dev())
    idx = torch.stack([torch.randperm(N - extra, generator=g)[:k] for _ in range(B)]).to(dev())
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(dev())
    b = (0.05 * torch.randn(D, generator=g)).to(dev())
    xo, yo = ops.gather_layernorm(x, idx, extra, w, b, 1e-6, out_dtype)
    ref_x = torch.cat([x[:, :extra], torch.gather(x[:, extra:], 1, idx[..., None].expand(-1, -1, D))], dim=1)
    assert torch.equal(xo, ref_x)                      # compaction is a pure copy: bit-exact
    ref_y = F.layer_norm(ref_x.double(), (D,), w.double(), b.double(), 1e-6)
    assert rel_err(yo.float(), ref_y) < (2e-6 if out_dtype == torch.float32 else 8e-3)
    xo2, none = ops.gather_layernorm(x, idx, extra, None, None, 1e-6, out_dtype)
    assert none is None and torch.equal(xo2, ref_x)


def test_fused_token_kernel():
    """EViT fused inattentive token (parity unpinned: upstream EViT semantics, restated in the oracle)."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(17)
    B, N, D, extra, R = 3, 514, 768, 2, 12
    k = 359
    x = torch.randn(B, N, D, generator=g).to(dev())
    partial = torch.rand(B, R, N, generator=g).to(dev())
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(dev())
    b = (0.05 * torch.randn(D, generator=g)).to(dev())
    score, idx, rest = ops.score_topk(partial, float(R), extra, k, want_rest=True)
    assert rest.dtype == torch.int32 and rest.shape == (B, N - extra - k)
    for c in range(B):   # kept and rest partition the candidates
        assert sorted(idx[c].tolist() + rest[c].tolist()) == list(range(N - extra))
    xo, yo = ops.gather_layernorm(x, idx, extra, w, b, 1e-6, torch.float32, score=score, rest_idx=rest)
    assert xo.shape == (B, extra + k + 1, D)
    from oracle import vit_oracle as vo
    ref = vo.gather_tokens(x.cpu(), idx.cpu(), extra, score.cpu(), fuse_token=True)
    assert torch.equal(xo[:, :extra + k].cpu(), ref[:, :extra + k])
    assert rel_err(xo[:, -1].cpu(), ref[:, -1].double()) < 2e-6
    assert rel_err(yo.cpu(), F.layer_norm(ref.double(), (D,), w.cpu().double(), b.cpu().double(), 1e-6)) < 5e-6


def test_pool_norm_both_variants():
    ops, _lib = _ops()
    g = torch.Generator().manual_seed(8)
    B, N, D = 5, 178, 768
    x = torch.randn(B, N, D, generator=g).to(dev())
    g1, b1, g2, b2 = [(torch.randn(D, generator=g) * 0.1 + (1 if i % 2 == 0 else 0)).to(dev()) for i in range(4)]
    out = ops.pool_norm(x, _lib.VARIANT_AUDIOMAE, g1, b1, 1e-6)
    ref = F.layer_norm(x[:, 1:].double().mean(1), (D,), g1.double(), b1.double(), 1e-6)
    assert rel_err(out, ref) < 2e-6
    out = ops.pool_norm(x, _lib.VARIANT_AST, g1, b1, 1e-6, g2, b2, 1e-5)
    t = F.layer_norm(x.double(), (D,), g1.double(), b1.double(), 1e-6)
    ref = F.layer_norm((t[:, 0] + t[:, 1]) / 2, (D,), g2.double(), b2.double(), 1e-5)
    assert rel_err(out, ref) < 2e-6


@pytest.mark.parametrize("B,C", [(64, 527), (3, 35), (1, 50)])
def test_head(B, C):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(9)
    pooled = torch.randn(B, 768, generator=g).to(dev())
    w = (torch.randn(C, 768, generator=g) * 0.02).to(dev())
    bias = torch.randn(C, generator=g).to(dev())
    out = ops.head(pooled, w, bias)
    assert rel_err(out, pooled.double() @ w.double().T + bias.double()) < 2e-6


def test_errors_surface_as_runtime_error():
    ops, _lib = _ops()
    with pytest.raises(RuntimeError, match="head dim|CUDA tensor|contiguous|libtpat"):
        ops.layernorm(torch.zeros(4, 100, device=dev()), torch.ones(100, device=dev()), torch.zeros(100, device=dev()),
                      1e-6, torch.float32)
