"""CPU, world_size 2, gloo: the N>1 host path (batch sharding + output gather, tpat/dist.py).
The model is a stand-in that returns deterministic per-clip outputs, so the test checks the
plumbing (ordering, ragged shards, int32 wire format) without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest  # noqa: F401


class _FakeModel:
    """logits[b] = clip id; topk_idx rows encode the clip id as well."""

    def __init__(self):
        self.last_topk_idx = None

    def __call__(self, x, keep_rate_list=None):
        ids = x[:, 0, 0].long()
        self.last_topk_idx = [None, (ids[:, None] * 10 + torch.arange(3)[None, :]).to(torch.int64), None]
        return ids[:, None].float().repeat(1, 4)


def _worker(rank, world, port, batch, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tpat import dist as tdist
        x = torch.arange(batch, dtype=torch.float32)[:, None, None].repeat(1, 2, 2)
        logits, idx = tdist.sharded_forward(_FakeModel(), x)
        good = logits.shape == (batch, 4) and torch.equal(logits[:, 0], torch.arange(batch).float())
        good &= idx[0] is None and idx[2] is None and idx[1].dtype == torch.int64
        good &= torch.equal(idx[1][:, 0], torch.arange(batch) * 10)
        s, e = tdist.shard_bounds(batch, world, rank)
        good &= tdist.shard_batch(x).shape[0] == e - s
        ok[rank] = bool(good)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_sharded_forward_gathers_in_global_order(batch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Manager().list([False, False])
    mp.spawn(_worker, args=(2, port, batch, ok), nprocs=2, join=True)
    assert list(ok) == [True, True]
