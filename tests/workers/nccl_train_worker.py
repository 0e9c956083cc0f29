"""torchrun worker (world_size 2, NCCL) for tests/test_gpu_50_multigpu.py: the fine-tune step batch-sharded over two
GPUs.  Each rank runs forward + backward on its half of a global batch; the engine all-reduces each backward stage's
gradient slice over NCCL from inside the backward (DDP semantics: mean).  Every rank then compares its gradients with
its own single-GPU backward of the WHOLE batch (BCE's mean over the batch makes the two equal), fp32 mode, 2e-5."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import test_gpu_60_train as t60
    import gpu_util
    gpu_util.dev = lambda: dev
    t60.dev = lambda: dev
    from oracle.golden_configs import GRAD_CONFIGS
    cfg = dict(GRAD_CONFIGS["audiomae_256_b2_train"])
    cfg["B"] = 4
    sd, x, y = t60.case_inputs(cfg)
    ok = True
    for precision, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        # single GPU, whole batch, no synchronisation
        full = t60.build_train_model(cfg, sd, precision, drop_path_rate=0.0)
        full._engines.get_train(dev).grad_sync = False
        F.binary_cross_entropy_with_logits(full(x.to(dev)), y.to(dev)).backward()
        ref = {k: p.grad.detach().clone() for k, p in full.named_parameters() if p.grad is not None}
        # sharded: 2 clips per rank, gradients averaged over NCCL inside the backward
        model = t60.build_train_model(cfg, sd, precision, drop_path_rate=0.0)
        s = slice(rank * 2, rank * 2 + 2)
        F.binary_cross_entropy_with_logits(model(x[s].to(dev)), y[s].to(dev)).backward()
        torch.cuda.synchronize()
        worst = 0.0
        for k, p in model.named_parameters():
            if k in ref:
                worst = max(worst, ((p.grad.double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)).item())
        if rank == 0:
            print(f"[nccl train {precision}] sharded + all-reduced vs single-GPU full batch: worst rel err {worst:.2e}", flush=True)
        ok &= worst < tol and len(ref) == 151
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_TRAIN_OK" if int(flag.item()) == 1 else "NCCL_TRAIN_MISMATCH", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
