"""torchrun worker (world_size 2, NCCL) for tests/test_gpu_50_multigpu.py: each rank runs the forward on its shard of
a global batch, the outputs are exchanged with tpat.dist's ONE packed all_gather, and every rank compares the result
with its own single-GPU forward of the whole global batch (bit-exact: the forward is batch-invariant)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from conftest import load_golden, make_case
    from tpat import dist as tdist
    import test_gpu_20_forward as t20
    t20.dev = lambda: dev
    g = load_golden("ast_spc2_b8_kr07")
    meta = g["meta"]
    sd, x = make_case(meta)
    x = torch.cat([x, x.flip(0)[:3]], dim=0).to(dev)          # 11 clips: ragged shards (6 + 5)
    model = t20.build_model(meta, sd, "bf16")
    with torch.no_grad():
        full = model(x)
        full_idx = [None if t is None else t.clone() for t in model.last_topk_idx]
        for _ in range(2):                                     # second call: cached buffers
            logits, idx = tdist.sharded_forward(model, x)
    ok = torch.equal(logits, full) and logits.shape[0] == 11
    for a, b in zip(idx, full_idx):
        ok &= (a is None) == (b is None)
        if a is not None:
            ok &= a.dtype == torch.int64 and torch.equal(a, b)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_GATHER_OK" if int(flag.item()) == 1 else "NCCL_GATHER_MISMATCH", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
