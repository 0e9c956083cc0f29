"""attention_tc4.cu (independent 32-key halves: own reference max / row sum / output accumulator per half, P written over
S in tensor memory, no partner exchange inside the key loop; TPAT_ATTN_V4=1) and attention_tc5.cu (the same kernel made
persistent: resident CTAs walk the (clip, head, tile) items; TPAT_ATTN_V5=1) against the fp64 softmax reference: every
tile / key-block / half tail shape, the lazy-rescale slow path (sum-triggered, including overflow to inf), the
log-sum-exp output, the AST tile offset, bit-equality of clips across batch positions, and -- for the persistent kernel --
batches large enough that every CTA walks several items."""
import os

import pytest
import torch

import conftest  # noqa: F401
from gpu_util import dev, ref_attention, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["v4", "v5"])
def _kernel(request, monkeypatch):
    monkeypatch.setenv("TPAT_ATTN_V4", "1")
    monkeypatch.setenv("TPAT_ATTN_V5", "1" if request.param == "v5" else "0")
    monkeypatch.setenv("TPAT_ATTN_V3", "0")
    return request.param


@pytest.mark.parametrize("N", [1, 2, 17, 32, 33, 48, 64, 65, 66, 97, 128, 129, 178, 200, 253, 256, 257, 360, 513, 514, 1025])
def test_attention_v4_matches_reference(N):
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(N)
    B, H = 3, 12
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    out, lse = ops.attention_train(qkv, B, N, H, min(1, N - 1), _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, attn = ref_attention(qkv, B, N, H, min(1, N - 1))
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((x[0] @ x[1].transpose(-2, -1)) * 0.125, dim=-1)
    e_o, e_l = rel_err(out.float(), ref_out), rel_err(lse, lse_ref)
    print(f"[attention v4] N={N}: out err {e_o:.2e}, lse err {e_l:.2e}")
    assert torch.isfinite(out.float()).all()
    assert e_o < 1e-2 and e_l < 1e-5
    v5 = os.environ["TPAT_ATTN_V5"]
    os.environ["TPAT_ATTN_V4"] = "0"; os.environ["TPAT_ATTN_V5"] = "0"
    old, _ = ops.attention(qkv, B, N, H, min(1, N - 1), _lib.SCORE_NONE, _lib.IMPL_TC)
    os.environ["TPAT_ATTN_V4"] = "1"; os.environ["TPAT_ATTN_V5"] = v5
    assert rel_err(out.float(), old.float()) < 1e-2
    again, _ = ops.attention(qkv[N:2 * N].contiguous(), 1, N, H, min(1, N - 1), _lib.SCORE_NONE, _lib.IMPL_TC)      # batch invariance
    assert torch.equal(again, out[N:2 * N])
    twice, _ = ops.attention(qkv, B, N, H, min(1, N - 1), _lib.SCORE_NONE, _lib.IMPL_TC)                           # run-to-run bits
    assert torch.equal(twice, out)


@pytest.mark.parametrize("N", [70, 200, 513, 514])
@pytest.mark.parametrize("gain", [24.0, 200.0])
def test_attention_v4_rescale_path(N, gain):
    """Keys whose scores grow with the key index force the lazily rescaled online softmax to raise its reference max and
    rescale the half's accumulator in tensor memory (several times per row); gain 200 drives exp2 to +inf before the
    redo, which the sum trigger must catch as well."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(22)
    B, H = 2, 12
    x = torch.randn(B, N, 3, H, 64, generator=g) * 1.5
    ramp = 1.0 + gain * torch.arange(N, dtype=torch.float32) / N
    x[:, :, 1] *= ramp[None, :, None, None]
    qkv = x.reshape(B * N, 3 * H * 64).to(dev()).to(torch.bfloat16)
    out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 1)
    xx = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((xx[0] @ xx[1].transpose(-2, -1)) * 0.125, dim=-1)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2
    assert rel_err(lse, lse_ref) < 1e-5


def test_attention_v4_one_dominant_late_key_in_the_second_half():
    """A single key in the SECOND half of a late block dominates every row: only that half rescales, the merge in the
    epilogue must weight the halves by their own reference maxima."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(3)
    B, H, N = 2, 12, 300
    x = torch.randn(B, N, 3, H, 64, generator=g)
    x[:, 250, 1] = 12.0 * torch.sign(torch.randn(B, H, 64, generator=g))      # large-norm key 250 (block 3, second half)
    x[:, :, 0] = x[:, :, 0].abs() * torch.sign(x[:, 250:251, 1])                                       # every query aligned with it
    qkv = x.reshape(B * N, 3 * H * 64).to(dev()).to(torch.bfloat16)
    out, _ = ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, attn = ref_attention(qkv, B, N, H, 1)
    assert attn[..., 250].min() > 0.99                  # the construction works: key 250 takes (almost) all the mass
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2


def test_attention_v4_ast_cls_row_split():
    """AST score blocks: tile 0 runs the two-pass kernel (cls row), tiles >= 1 the v4 kernel with a tile offset."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(5)
    B, H, N = 2, 12, 514
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    out, partial = ops.attention(qkv, B, N, H, 2, _lib.SCORE_CLS_ROW, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 2)
    assert rel_err(out.float(), ref_out) < 1e-2


@pytest.mark.parametrize("B,N", [(32, 66), (32, 129), (24, 178), (16, 360), (16, 513), (40, 200)])
def test_attention_many_items_per_cta(B, N):
    """More (clip, head, tile) items than resident CTAs (2 x 148): the persistent kernel walks several items per CTA --
    Q double buffer, K / V ring and S / P parities running across items, accumulator hand-over (o_empty), deferred
    store retirement.  v4 and v5 must agree BIT FOR BIT (same per-row arithmetic), and both match the reference."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(B * 1000 + N)
    H = 12
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 1)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((x[0] @ x[1].transpose(-2, -1)) * 0.125, dim=-1)
    assert rel_err(lse, lse_ref) < 1e-5
    v5 = os.environ["TPAT_ATTN_V5"]
    os.environ["TPAT_ATTN_V5"] = "0"
    out4, lse4 = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    os.environ["TPAT_ATTN_V5"] = v5
    assert torch.equal(out, out4) and torch.equal(lse, lse4)
    one, _ = ops.attention(qkv[5 * N:6 * N].contiguous(), 1, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)      # batch invariance
    assert torch.equal(one, out[5 * N:6 * N])


def test_attention_many_items_rescale_path():
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(23)
    B, H, N = 32, 12, 200
    x = torch.randn(B, N, 3, H, 64, generator=g) * 1.5
    ramp = 1.0 + 200.0 * torch.arange(N, dtype=torch.float32) / N
    x[:, :, 1] *= ramp[None, :, None, None]
    qkv = x.reshape(B * N, 3 * H * 64).to(dev()).to(torch.bfloat16)
    out, _ = ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 1)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2
