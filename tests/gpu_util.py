"""Helpers shared by the -m gpu tests: plain PyTorch references (fp32 / fp64) for single kernels."""
import torch
import torch.nn.functional as F


def dev():
    return torch.device("cuda:0")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (normalised by the reference's largest magnitude)."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def ref_attention(qkv: torch.Tensor, B: int, N: int, H: int, num_extra: int, p_dtype=None):
    """qkv [B*N, 3*H*64] (any float dtype) -> (out [B*N, H*64] fp64, P [B,H,N,N] fp64).
    Follows models_vit.py:75-95.  ``p_dtype``: round P to that dtype before P @ V (bf16 kernels)."""
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    attn = ((q @ k.transpose(-2, -1)) * (64 ** -0.5)).softmax(dim=-1)
    pa = attn if p_dtype is None else attn.to(p_dtype).double()
    out = (pa @ v).transpose(1, 2).reshape(B * N, H * 64)
    return out, attn


def ref_score(attn: torch.Tensor, num_extra: int, mode: str) -> torch.Tensor:
    if mode == "colmean":      # models_vit.py:113
        return attn[:, :, num_extra:, num_extra:].mean(dim=(1, 2))
    return attn[:, :, 0, num_extra:].mean(dim=1)   # ast_models.py:124


def set_overlap(a: torch.Tensor, b: torch.Tensor) -> float:
    tot = 0.0
    for ra, rb in zip(a.tolist(), b.tolist()):
        tot += len(set(ra) & set(rb)) / max(1, len(ra))
    return tot / a.shape[0]
