"""Multi-GPU paths on real hardware (skipped on a one-GPU box): the north-star eval collective over NCCL
(engine_finetune.py:246-248 -> util/stat.py:12-22) and the reference's nn.DataParallel wrapper (traintest.py:79,286)."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

import conftest  # noqa: F401
from conftest import load_golden, make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "workers", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


def test_packed_all_gather_over_nccl_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = _torchrun("nccl_gather_worker.py", 2)
    assert r.returncode == 0 and "NCCL_GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_data_parallel_replicas_have_their_own_engine():
    """nn.DataParallel (the reference's AST wrapper): replicas share the model object's EnginePool, and each replica
    thread must pack / run on ITS device.  Output must equal the single-GPU forward bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import test_gpu_20_forward as t20
    g = load_golden("ast_spc2_b8_kr07")
    meta = g["meta"]
    sd, x = make_case(meta)
    model = t20.build_model(meta, sd, "bf16")
    with torch.no_grad():
        single = model(x.to("cuda:0"))
        dp = nn.DataParallel(model, device_ids=[0, 1])
        for _ in range(2):
            out = dp(x.to("cuda:0"))
    assert out.device.index == 0 and torch.equal(out, single)
    assert sorted(model._engines._engines) == [0, 1]


def test_train_step_grad_allreduce_over_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not os.path.exists(os.path.join(ROOT, "tests", "workers", "nccl_train_worker.py")):
        pytest.skip("train worker not present")
    r = _torchrun("nccl_train_worker.py", 2)
    assert r.returncode == 0 and "NCCL_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
