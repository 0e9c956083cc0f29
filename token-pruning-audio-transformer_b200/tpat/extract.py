"""Host-side glue for the feature-extraction output contract.

The reference's analysis pipeline (audiomae/extract_stats.py) consumes what its eval loop saves from
``model(x, flag_extract_features=True)``: one ``{key}.{idx:04d}.pth`` file per feature-dict entry
(engine_finetune.py:189-193), block-local ``topk_idx`` composed into mel-patch coordinates
(util/token_reduction_utils.py:8-19) and spectrograms masked to the kept patches (util/misc.py:422-437).
These helpers reproduce that contract for the feature dict the tpat models return; they are bookkeeping on
small CPU index tensors, not part of the GPU hot path.
"""
import os
from typing import Dict, List, Sequence

import torch


def save_feature_dict(feature_dict: Dict[str, torch.Tensor], extract_features_path: str, idx: int) -> List[str]:
    """Write every entry as ``{extract_features_path}/{key}.{idx:04d}.pth`` (engine_finetune.py:189-193)."""
    os.makedirs(extract_features_path, exist_ok=True)
    paths = []
    for key, value in feature_dict.items():
        path = f"{extract_features_path}/{key}.{idx:04d}.pth"
        torch.save(value, path)
        paths.append(path)
    return paths


def get_melspec_idx(idxs: Sequence[torch.Tensor], fuse_token: bool = False) -> List[torch.Tensor]:
    """Compose block-local top-k indices into mel-patch coordinates: ``idxs[i] = gather(idxs[i-1], 1, idxs[i])``
    (util/token_reduction_utils.py:8-19).  With ``fuse_token`` the previous level gets one extra column (index 0)
    standing for the fused token, exactly as the reference helper does."""
    out = [t.clone() for t in idxs]
    for i in range(1, len(out)):
        tmp = out[i - 1]
        if fuse_token:
            tmp = torch.cat([tmp, torch.zeros(tmp.size(0), 1, dtype=tmp.dtype, device=tmp.device)], dim=1)
        out[i] = torch.gather(tmp, dim=1, index=out[i])
    return out


def topk_indices_of(feature_dict: Dict[str, torch.Tensor]) -> List[torch.Tensor]:
    """The ``block-i.topk_idx`` entries of a feature dict in block order."""
    keys = sorted((k for k in feature_dict if k.endswith(".topk_idx")), key=lambda k: int(k.split(".")[0].split("-")[1]))
    return [feature_dict[k] for k in keys]


def apply_mask(x: torch.Tensor, idx: torch.Tensor, patch_size: int = 16) -> torch.Tensor:
    """Keep only the patches listed in ``idx`` (mel-patch coordinates), zero the rest (util/misc.py:422-437).
    x [B, C, H, W]; idx [B, T] with values in [0, (H/p)*(W/p))."""
    B, C, H, W = x.shape
    h, w = H // patch_size, W // patch_size
    patches = x.reshape(B, C, h, patch_size, w, patch_size).permute(0, 1, 3, 5, 2, 4).reshape(B, C * patch_size * patch_size, h * w)
    out = torch.zeros_like(patches)
    gidx = idx.unsqueeze(1).expand(-1, patches.size(1), -1)
    out.scatter_(2, gidx, torch.gather(patches, 2, gidx))
    return out.reshape(B, C, patch_size, patch_size, h, w).permute(0, 1, 4, 2, 5, 3).reshape(B, C, H, W)
