"""attention_tc3.cu (two query tiles per CTA, one thread per row, 128-key blocks; selected with TPAT_ATTN_V3=1 or by
default once enabled) against the fp64 softmax reference: all tile / key-block tail shapes, the lazy-rescale slow path,
the log-sum-exp output, and bit-equality of clips across batch positions."""
import os

import pytest
import torch

import conftest  # noqa: F401
from gpu_util import dev, ref_attention, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _v3(monkeypatch):
    monkeypatch.setenv("TPAT_ATTN_V3", "1")


@pytest.mark.parametrize("N", [2, 17, 66, 128, 129, 200, 256, 257, 360, 513, 514, 1025])
def test_attention_v3_matches_reference(N):
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(N)
    B, H = 3, 12
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, attn = ref_attention(qkv, B, N, H, 1)
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((x[0] @ x[1].transpose(-2, -1)) * 0.125, dim=-1)
    e_o, e_l = rel_err(out.float(), ref_out), rel_err(lse, lse_ref)
    print(f"[attention v3] N={N}: out err {e_o:.2e}, lse err {e_l:.2e}")
    assert torch.isfinite(out.float()).all()
    assert e_o < 1e-2 and e_l < 1e-5
    os.environ["TPAT_ATTN_V3"] = "0"
    old, _ = ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    os.environ["TPAT_ATTN_V3"] = "1"
    assert rel_err(out.float(), old.float()) < 1e-2
    again, _ = ops.attention(qkv[N:2 * N].contiguous(), 1, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)      # batch invariance
    assert torch.equal(again, out[N:2 * N])


@pytest.mark.parametrize("N", [200, 513, 514])
def test_attention_v3_rescale_path(N):
    """Keys whose scores grow with the key index force the lazily rescaled online softmax to raise its reference max and
    rescale O in tensor memory (several times per row)."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(22)
    B, H = 2, 12
    x = torch.randn(B, N, 3, H, 64, generator=g) * 1.5
    ramp = 1.0 + 24.0 * torch.arange(N, dtype=torch.float32) / N
    x[:, :, 1] *= ramp[None, :, None, None]
    qkv = x.reshape(B * N, 3 * H * 64).to(dev()).to(torch.bfloat16)
    out, _ = ops.attention(qkv, B, N, H, 1, _lib.SCORE_NONE, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 1)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), ref_out) < 1e-2


def test_attention_v3_ast_cls_row_split():
    """AST score blocks: tile 0 runs the two-pass kernel (cls row), tiles >= 1 the v3 kernel with a tile offset."""
    from tpat import ops, _lib
    g = torch.Generator().manual_seed(5)
    B, H, N = 2, 12, 514
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).to(dev()).to(torch.bfloat16)
    out, partial = ops.attention(qkv, B, N, H, 2, _lib.SCORE_CLS_ROW, _lib.IMPL_TC)
    ref_out, _ = ref_attention(qkv, B, N, H, 2)
    assert rel_err(out.float(), ref_out) < 1e-2
