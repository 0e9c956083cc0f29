"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Clips are independent (top-k is per clip), so the forward shards by batch with the weights
replicated and NO collective on the data path.  The only exchange is the
eval-time gather of outputs, mirroring ``concat_all_gather`` in the reference
(audiomae/util/stat.py:12-22, used at engine_finetune.py:246-248): logits plus the kept top-k
indices.  Backend: "nccl" on GPUs (NVLink 5 / NVSwitch), "gloo" in the CPU tests.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) slice of ``n_items`` for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, world_size: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    """This rank's slice of a global batch along dim 0."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    s, e = shard_bounds(x.shape[0], world_size, rank)
    return x[s:e]


def _gather_ragged(t: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """all_gather along dim 0 of per-rank tensors whose dim-0 sizes are ``counts`` (known on every rank)."""
    world = dist.get_world_size()
    mx = max(counts)
    if t.shape[0] < mx:  # pad the short shards so every rank contributes the same shape
        pad = torch.zeros((mx - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def gather_outputs(logits: torch.Tensor, topk_idx: Sequence[Optional[torch.Tensor]], global_batch: int
                   ) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
    """Gather the sharded forward outputs on every rank, in global clip order.

    ``topk_idx`` is the per-block list a forward returns (None where a block does not prune).
    Indices travel as int32 (<= 4096 tokens) and are widened back to int64 at the API boundary.
    """
    world = dist.get_world_size()
    counts = [shard_bounds(global_batch, world, r)[1] - shard_bounds(global_batch, world, r)[0] for r in range(world)]
    all_logits = _gather_ragged(logits, counts)
    all_idx: List[Optional[torch.Tensor]] = []
    for t in topk_idx:
        all_idx.append(None if t is None else _gather_ragged(t.to(torch.int32), counts).to(torch.int64))
    return all_logits, all_idx


def sharded_forward(model, x_global: torch.Tensor, keep_rate_list=None) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
    """Run ``model`` on this rank's shard of ``x_global`` and gather (logits, topk_idx lists).
    The model must expose ``last_topk_idx`` after a forward (both tpat model classes do)."""
    shard = shard_batch(x_global)
    logits = model(shard, keep_rate_list=keep_rate_list)
    return gather_outputs(logits, model.last_topk_idx, x_global.shape[0])
