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


class PackedGather:
    """ONE collective for the eval-time gather: logits (fp32) and the kept-index blocks (int32 on the wire, <= 4096
    tokens) of this rank are packed into one pre-allocated int32 row buffer [rows_max, width] and exchanged with a single
    ``all_gather_into_tensor`` -- the payload is <= 1 MB, i.e. latency-bound, so one launch instead of 1 + n_prune
    list-based ``all_gather`` calls (+ pad / cat kernels) is what matters.  Buffers are cached per (shape, device), so a
    steady-state step issues two small pack copies per tensor, the collective, and views for the unpack.

    Mirrors ``concat_all_gather`` (audiomae/util/stat.py:12-22 at engine_finetune.py:246-248)."""

    def __init__(self):
        self._buf = {}

    def _buffers(self, rows_max, width, world, device):
        key = (rows_max, width, world, device)
        ent = self._buf.get(key)
        if ent is None:
            send = torch.zeros(rows_max, width, dtype=torch.int32, device=device)
            recv = torch.empty(world * rows_max, width, dtype=torch.int32, device=device)
            ent = self._buf[key] = (send, recv)
        return ent

    def __call__(self, logits: torch.Tensor, topk_idx: Sequence[Optional[torch.Tensor]], global_batch: int
                 ) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
        world = dist.get_world_size()
        counts = [shard_bounds(global_batch, world, r)[1] - shard_bounds(global_batch, world, r)[0] for r in range(world)]
        rows_max = max(counts)
        n_local, C = logits.shape
        widths = [C] + [0 if t is None else t.shape[1] for t in topk_idx]
        send, recv = self._buffers(rows_max, sum(widths), world, logits.device)
        send[:n_local, :C].copy_(logits.detach().float().view(torch.int32) if logits.dtype == torch.float32
                                 else logits.detach().float().contiguous().view(torch.int32))
        off = C
        for t, w in zip(topk_idx, widths[1:]):
            if t is not None:
                send[:n_local, off:off + w].copy_(t)            # int64 -> int32 in the copy kernel
            off += w
        dist.all_gather_into_tensor(recv, send)
        rv = recv.view(world, rows_max, -1)
        if all(c == rows_max for c in counts):
            flat = rv.reshape(world * rows_max, -1)
        else:
            flat = torch.cat([rv[r, :c] for r, c in enumerate(counts)], dim=0)
        all_logits = flat[:, :C].contiguous().view(torch.float32)
        all_idx: List[Optional[torch.Tensor]] = []
        off = C
        for t, w in zip(topk_idx, widths[1:]):
            all_idx.append(None if t is None else flat[:, off:off + w].to(torch.int64))
            off += w
        return all_logits, all_idx


_packed_gather = PackedGather()


def gather_outputs(logits: torch.Tensor, topk_idx: Sequence[Optional[torch.Tensor]], global_batch: int
                   ) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
    """Gather the sharded forward outputs on every rank, in global clip order, with ONE packed collective.

    ``topk_idx`` is the per-block list a forward returns (None where a block does not prune).
    Indices travel as int32 and are widened back to int64 at the API boundary."""
    return _packed_gather(logits, topk_idx, global_batch)


def sharded_forward(model, x_global: torch.Tensor, keep_rate_list=None) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
    """Run ``model`` on this rank's shard of ``x_global`` and gather (logits, topk_idx lists).
    The model must expose ``last_topk_idx`` after a forward (both tpat model classes do)."""
    shard = shard_batch(x_global)
    logits = model(shard, keep_rate_list=keep_rate_list)
    return gather_outputs(logits, model.last_topk_idx, x_global.shape[0])
