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


def _grad_sync_worker(rank, world, port, ok):
    """The bucketed gradient all-reduce of the fine-tune step (TrainEngine.allreduce_stage / finish_sync) on CPU tensors
    over gloo: every backward stage's gradients are ONE contiguous slice of the flat buffer, reduced in place."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn as nn
        from tpat import models_vit, _lib
        from tpat.train import TrainEngine
        torch.manual_seed(0)
        m = models_vit.vit_base_patch16(num_classes=10, drop_path_rate=0.0, mean_pooling=True, mask_2d=True, target_length=128,
                                        drop_loc=(3, 6, 9), base_keep_rate=0.7)
        m.patch_embed = models_vit.PatchEmbed((128, 128), 16, 1, 768)
        m.pos_embed = nn.Parameter(torch.zeros(1, 65, 768), requires_grad=False)
        eng = TrainEngine(_lib.VARIANT_AUDIOMAE, 12, 768, 12, 3072)
        eng.attach(m._train_entries(m._train_roles()), {})
        for _, _, p in eng.entries:                      # rank-dependent "gradients", written through the views
            p.grad = eng.grad_view(p)
            p.grad.fill_(float(rank + 1))
        m.head.bias.grad[3] = 10.0 * (rank + 1)
        world_seen = eng.sync_world()
        works = [w for w in (eng.allreduce_stage(s, world_seen) for s in range(13, -1, -1)) if w is not None]
        eng.finish_sync(works, world_seen)
        good = world_seen == 2 and len(works) == 14
        good &= all(bool(torch.all(p.grad.flatten()[:2] == 1.5)) for _, _, p in eng.entries if p is not m.head.bias)
        good &= abs(m.head.bias.grad[3].item() - 15.0) < 1e-6 and abs(m.head.bias.grad[0].item() - 1.5) < 1e-6
        eng.grad_sync = False                            # accumulation steps: no collective
        good &= eng.sync_world() == 1 and eng.allreduce_stage(5, eng.sync_world()) is None
        ok[rank] = bool(good)
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_allreduce_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Manager().list([False, False])
    mp.spawn(_grad_sync_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [True, True]
