"""Fine-tune step: host-side driver of ``tpat_train_forward`` / ``tpat_train_backward``.

The reference fine-tunes with plain PyTorch autograd (audiomae/engine_finetune.py:102-105 forward + loss,
util/misc.py:259-273 ``loss.backward()`` + optimizer step, main_finetune.py:459-461 DDP gradient all-reduce).
Here ``model(x, keep_rate_list)`` in training mode returns logits that carry ONE autograd node: its backward runs
the native backward of the whole network (every kernel in libtpat.so) and leaves the gradients in ``p.grad`` of
every parameter, so ``criterion(model(x), y).backward(); optimizer.step()`` works unchanged.

Memory layout (B200-first): all trainable parameters live in ONE flat fp32 buffer (``p.data`` are views into it),
all gradients in a second one (``p.grad`` are views), ordered by backward stage -- classifier head, blocks, patch
embedding -- so that the gradients a stage produces are one contiguous slice: the DDP-style bucketed all-reduce is
an in-place NCCL all-reduce of that slice, launched right after the stage's kernels are enqueued and overlapping
the next stage's compute (no bucket copies).  The tcgen05 path also keeps a flat bf16 copy of the parameters (GEMM
operands) and [in, out] copies of the four matrices of each block (data gradients dX = dY W).

Gradients are written straight into ``p.grad`` (the autograd node returns ``None`` for its anchor input); a
backward after ``optimizer.zero_grad()`` overwrites, a second backward without it accumulates -- autograd's contract.
"""
import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import TrainArgs, check, lib
from .engine import pruning_schedule, tokens_entering

BLOCK_ROLES = _lib.BLOCK_GRAD_NAMES   # ln1_g ... fc2_b
MATRIX_ROLES = ("qkv_w", "proj_w", "fc1_w", "fc2_w")
_DGRAD_TRANSPOSE = os.environ.get("TPAT_DGRAD_TRANSPOSE") is not None     # A/B switch: [in, out] bf16 copies instead of w_kn


class TrainEngine:
    """Training state of one model on one device.  ``entries`` = ordered (stage, role, parameter) triples from the
    model (``model._train_entries()``): stage depth + 1 = head / final norm, i + 1 = block i, 0 = patch embedding."""

    def __init__(self, variant: int, depth: int, D: int, H: int, Dh: int):
        self.variant, self.depth, self.D, self.H, self.Dh = variant, depth, D, H, Dh
        self.num_extra = 2 if variant == _lib.VARIANT_AST else 1
        self.flat_p: Optional[torch.Tensor] = None
        self.flat_g: Optional[torch.Tensor] = None
        self.flat_pb: Optional[torch.Tensor] = None          # bf16 operand copy (tcgen05 path)
        self.entries: List[tuple] = []
        self.offsets: Dict[int, Tuple[int, int]] = {}        # id(param) -> (offset, numel)
        self.stage_slices: Dict[int, Tuple[int, int]] = {}   # stage -> (offset, numel) of its gradients in flat_g
        self._wt: Dict[tuple, torch.Tensor] = {}
        self._versions = None
        self._operands_fresh = False
        self._frozen: Dict[str, torch.Tensor] = {}
        self._cache: Dict[tuple, dict] = {}
        self.grad_sync = True            # all-reduce gradients across torch.distributed ranks (DDP semantics: mean)
        self.fused_grad_scale = False    # FusedAdamW sets it: the 1 / world averaging happens inside its kernel
        self.pending_grad_scale = 1.0
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self.last_launches = 0

    # ---- flat parameter / gradient buffers ---------------------------------------------------
    def attach(self, entries: Sequence[tuple], frozen: Dict[str, torch.Tensor]) -> None:
        """(Re)build the flat buffers when the model's parameter tensors are not (all) views of them any more
        (first call, ``model.to(device)``, ``load_state_dict`` keeps the views so it does not trigger this)."""
        trainable = [(st, role, p) for st, role, p in entries if p.requires_grad]
        ok = self.flat_p is not None and len(trainable) == len(self.entries)
        if ok:
            base, end = self.flat_p.data_ptr(), self.flat_p.data_ptr() + self.flat_p.numel() * 4
            for (st, role, p), (st0, role0, p0) in zip(trainable, self.entries):
                off, n = self.offsets.get(id(p), (-1, 0))
                if p is not p0 or off < 0 or p.data_ptr() != base + off * 4 or n != p.numel() or p.dtype != torch.float32:
                    ok = False
                    break
        self._frozen = frozen
        if ok:
            return
        dev = trainable[0][2].device
        # every tensor starts on a 64-element (256-byte fp32 / 128-byte bf16) boundary: TMA and vector-load alignment
        total, layout = 0, []
        for st, role, p in trainable:
            layout.append(total)
            total += (p.numel() + 63) // 64 * 64
        flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.offsets, self.stage_slices = {}, {}
        with torch.no_grad():
            for (st, role, p), off in zip(trainable, layout):
                n = p.numel()
                flat_p[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = flat_p[off:off + n].view(p.shape)
                p.grad = None
                self.offsets[id(p)] = (off, n)
                lo, cnt = self.stage_slices.get(st, (off, 0))
                self.stage_slices[st] = (lo, off + (n + 63) // 64 * 64 - lo)
        self.flat_p, self.flat_g, self.entries = flat_p, flat_g, trainable
        self.flat_pb = None
        self._wt, self._cache = {}, {}
        self._versions, self._operands_fresh = None, False

    def grad_view(self, p: torch.Tensor) -> torch.Tensor:
        off, n = self.offsets[id(p)]
        return self.flat_g[off:off + n].view(p.shape)

    def _operand(self, p: torch.Tensor, impl: int) -> torch.Tensor:
        """The GEMM-operand copy of a matrix parameter: the fp32 master itself (fp32 path) or its bf16 copy."""
        if impl == _lib.IMPL_SIMT:
            return p.detach()
        off, n = self.offsets[id(p)]
        return self.flat_pb[off:off + n].view(p.shape)

    def mark_updated(self, bf16_fresh: bool = False) -> None:
        """Called by ``FusedAdamW.step`` (which writes the flat buffers without bumping tensor versions)."""
        self._versions = None
        self._operands_fresh = bf16_fresh

    def refresh_operands(self, impl: int) -> None:
        """bf16 copy of the flat parameters (tcgen05 path: the data-gradient GEMMs read the forward weights as MN-major
        operands, no transposed copies) or fp32 [in, out] copies of the block matrices (fp32 path), rebuilt when the
        weights changed."""
        versions = tuple(p._version for _, _, p in self.entries)
        if versions == self._versions and (impl == _lib.IMPL_SIMT or self.flat_pb is not None) and self._wt.get("impl") == impl:
            return
        stream = torch.cuda.current_stream().cuda_stream
        if impl == _lib.IMPL_TC:
            if self.flat_pb is None:
                self.flat_pb = torch.empty(self.flat_p.numel(), device=self.flat_p.device, dtype=torch.bfloat16)
                self._operands_fresh = False
            if not self._operands_fresh:
                self.flat_pb.copy_(self.flat_p)            # (FusedAdamW refreshes it inside its own kernel)
        act = torch.bfloat16 if impl == _lib.IMPL_TC else torch.float32
        dt = _lib.BF16 if impl == _lib.IMPL_TC else _lib.F32
        for st, role, p in self.entries:
            if role in MATRIX_ROLES and st >= 1 and (impl == _lib.IMPL_SIMT or _DGRAD_TRANSPOSE):
                key = (st, role)
                out_f, in_f = p.shape
                t = self._wt.get(key)
                if t is None or t.dtype != act:
                    t = self._wt[key] = torch.empty(in_f, out_f, device=p.device, dtype=act)
                check(lib.tpat_transpose(p.data_ptr(), _lib.F32, in_f, t.data_ptr(), dt, out_f, out_f, in_f, stream), "tpat_transpose")
        self._wt["impl"] = impl
        self._versions, self._operands_fresh = versions, False

    # ---- one step -------------------------------------------------------------------------
    def _role_ptr(self, role_map, role, impl, operand=False):
        p = role_map.get(role)
        if p is None:
            return None
        if operand:
            return self._operand(p, impl).data_ptr() if p.requires_grad else self._frozen_operand(p, impl).data_ptr()
        return p.data_ptr()

    def _frozen_operand(self, p, impl):
        if impl == _lib.IMPL_SIMT:
            return p.detach()
        key = ("frozen", id(p), p._version)
        t = self._wt.get(key)
        if t is None:
            t = self._wt[key] = p.detach().to(torch.bfloat16).contiguous()
        return t

    def _build(self, spec, prune, keep, impl, num_classes, want_all_scores, n_keep, roles):
        """TrainArgs + buffers for one (shape, schedule); cached."""
        B, T, F = spec.shape
        dev = spec.device
        a = TrainArgs()
        f = a.fwd
        f.variant, f.impl = self.variant, impl
        f.B, f.T, f.F = B, T, F
        f.depth, f.D, f.H, f.Dh, f.num_classes = self.depth, self.D, self.H, self.Dh, num_classes
        for i in range(self.depth):
            f.prune[i], f.keep[i] = prune[i], keep[i]
        f.want_all_scores = 1 if want_all_scores else 0
        f.ln_eps, f.norm_eps, f.head_ln_eps = 1e-6, 1e-6, 1e-5
        n0 = n_keep if n_keep else (T // 16) * (F // 16)
        entering = tokens_entering(n0, prune, keep, False)
        scores = [torch.empty(B, entering[i], device=dev, dtype=torch.float32) if (prune[i] or want_all_scores) else None
                  for i in range(self.depth)]
        idxs = [torch.empty(B, keep[i], device=dev, dtype=torch.int64) if prune[i] else None for i in range(self.depth)]
        for i in range(self.depth):
            f.scores[i] = scores[i].data_ptr() if scores[i] is not None else None
            f.topk_idx[i] = idxs[i].data_ptr() if idxs[i] is not None else None
        a.n_keep = n_keep or 0
        ent = {"args": a, "scores": scores, "idxs": idxs}
        a.mask_keep_idx = 1 if n_keep else None          # placeholder so that the size queries see the masked schedule
        self._fill_weights(a, roles, impl)
        need_s, need_w = lib.tpat_train_saved_bytes(ctypes.byref(a)), lib.tpat_train_bwd_workspace_bytes(ctypes.byref(a))
        if need_s == 0 or need_w == 0:
            raise RuntimeError(f"libtpat tpat_train_*_bytes failed: {_lib.last_error()}")
        ent["saved"] = torch.empty(need_s, device=dev, dtype=torch.uint8)
        ent["bwd_ws"] = torch.empty(need_w, device=dev, dtype=torch.uint8)
        a.saved, a.saved_bytes = ent["saved"].data_ptr(), need_s
        a.bwd_workspace, a.bwd_workspace_bytes = ent["bwd_ws"].data_ptr(), need_w
        a.mask_keep_idx = None
        return ent

    def _fill_weights(self, a: TrainArgs, roles, impl) -> None:
        f = a.fwd
        top, blocks = roles
        op = lambda p: (self._operand(p, impl) if p.requires_grad else self._frozen_operand(p, impl))
        f.patch_w = op(top["patch_w"]).data_ptr()
        f.patch_b = top["patch_b"].data_ptr()
        f.extra_tok = top["extra_tok"].data_ptr()
        f.pos = top["pos"].data_ptr()
        f.norm_g, f.norm_b = top["norm_g"].data_ptr(), top["norm_b"].data_ptr()
        if top.get("head_ln_g") is not None:
            f.head_ln_g, f.head_ln_b = top["head_ln_g"].data_ptr(), top["head_ln_b"].data_ptr()
        f.head_w, f.head_b = top["head_w"].data_ptr(), top["head_b"].data_ptr()
        gptr = lambda p: (self.grad_view(p).data_ptr() if (p is not None and p.requires_grad) else None)
        a.d_patch_w, a.d_patch_b = gptr(top["patch_w"]), gptr(top["patch_b"])
        a.d_extra_tok = gptr(top["extra_tok_first"])
        a.d_pos = gptr(top["pos"])
        a.d_norm_g, a.d_norm_b = gptr(top["norm_g"]), gptr(top["norm_b"])
        a.d_head_ln_g, a.d_head_ln_b = gptr(top.get("head_ln_g")), gptr(top.get("head_ln_b"))
        a.d_head_w, a.d_head_b = gptr(top["head_w"]), gptr(top["head_b"])
        for i, blk in enumerate(blocks):
            bw, gr, wt = f.blocks[i], a.grads[i], a.wt[i]
            for role in BLOCK_ROLES:
                p = blk[role]
                setattr(bw, role, (op(p) if role in MATRIX_ROLES else p).data_ptr())
                setattr(gr, role, gptr(p))
            for role in MATRIX_ROLES:
                t = self._wt.get((i + 1, role))
                setattr(wt, role + "t", t.data_ptr() if t is not None and t.dtype == (torch.float32 if impl == _lib.IMPL_SIMT else torch.bfloat16) else None)

    def forward(self, spec: torch.Tensor, keep_rates: Sequence[float], num_classes: int, precision: str, roles,
                drop_scales: Optional[List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]] = None,
                mask_keep_idx: Optional[torch.Tensor] = None, want_all_scores: bool = False):
        """Training-mode forward.  Returns (logits, scores, idxs, ctx); ``ctx`` is what ``backward`` needs."""
        if precision == "bf16+score32":
            raise NotImplementedError("bf16+score32 is an inference precision mode")
        if not spec.is_cuda:
            raise RuntimeError("tpat: input must be a CUDA tensor; there is no CPU path")
        if not lib.tpat_device_ok():
            raise RuntimeError("tpat: the current device is not compute capability 10.x (B200, sm_100a)")
        impl = _lib.IMPL_SIMT if precision == "fp32" else _lib.IMPL_TC
        if spec.dtype != torch.float32 or not spec.is_contiguous():
            spec = spec.float().contiguous()
        B, T, F = spec.shape
        n_keep = int(mask_keep_idx.shape[1]) if mask_keep_idx is not None else 0
        n0 = n_keep if n_keep else (T // 16) * (F // 16)
        prune, keep = pruning_schedule(n0, self.num_extra, keep_rates, False)
        self.refresh_operands(impl)
        key = (B, T, F, tuple(prune), tuple(keep), impl, num_classes, bool(want_all_scores), n_keep)
        ent = self._cache.get(key)
        if ent is None:
            if len(self._cache) >= 4:                      # each entry owns an activation arena: keep few
                self._cache.pop(next(iter(self._cache)))
            ent = self._cache[key] = self._build(spec, prune, keep, impl, num_classes, want_all_scores, n_keep, roles)
        a: TrainArgs = ent["args"]
        self._fill_weights(a, roles, impl)                 # pointers may move when buffers are rebuilt; cheap
        logits = torch.empty(B, num_classes, device=spec.device, dtype=torch.float32)
        a.fwd.spec, a.fwd.logits = spec.data_ptr(), logits.data_ptr()
        keepalive = [spec, logits]
        for i in range(self.depth):
            for j in (0, 1):
                t = drop_scales[i][j] if drop_scales is not None else None
                a.drop_scale[i][j] = t.data_ptr() if t is not None else None
                keepalive.append(t)
        if mask_keep_idx is not None:
            mask_keep_idx = mask_keep_idx.contiguous()
            a.mask_keep_idx = mask_keep_idx.data_ptr()
            keepalive.append(mask_keep_idx)
        else:
            a.mask_keep_idx = None
        check(lib.tpat_train_forward(ctypes.byref(a), torch.cuda.current_stream().cuda_stream), "tpat_train_forward")
        ent["keepalive"] = keepalive
        return logits, ent["scores"], ent["idxs"], ent

    def backward(self, ent: dict, dlogits: torch.Tensor) -> None:
        a: TrainArgs = ent["args"]
        dlogits = dlogits.contiguous().float()
        a.dlogits = dlogits.data_ptr()
        fresh = all(p.grad is None for _, _, p in self.entries)
        if fresh:
            self.flat_g.zero_()
        for _, _, p in self.entries:
            if p.grad is None:
                p.grad = self.grad_view(p)
        stream = torch.cuda.current_stream()
        world = self.sync_world()
        works = []
        for stage in range(self.depth + 1, -1, -1):
            check(lib.tpat_train_backward(ctypes.byref(a), stage, stage, stream.cuda_stream), "tpat_train_backward")
            w = self.allreduce_stage(stage, world)
            if w is not None:
                works.append(w)
        self.finish_sync(works, world)
        ent["dlogits"] = dlogits


    # ---- gradient synchronisation (DDP-equivalent, main_finetune.py:459-461) -------------------------
    def sync_world(self) -> int:
        return dist.get_world_size() if (self.grad_sync and dist.is_available() and dist.is_initialized()) else 1

    def allreduce_stage(self, stage: int, world: int):
        """In-place SUM all-reduce of the contiguous gradient slice of one backward stage (the bucket), asynchronous:
        the collective waits for the kernels enqueued so far and overlaps the next stage's compute."""
        if world <= 1 or stage not in self.stage_slices:
            return None
        off, n = self.stage_slices[stage]
        return dist.all_reduce(self.flat_g[off:off + n], op=dist.ReduceOp.SUM, async_op=True)

    def finish_sync(self, works, world: int) -> None:
        for w in works:
            w.wait()
        if world > 1:
            if self.fused_grad_scale:
                self.pending_grad_scale = 1.0 / world    # DDP averages: folded into the FusedAdamW kernel
            else:
                self.flat_g.mul_(1.0 / world)


class _TrainFn(torch.autograd.Function):
    """One autograd node for the whole network.  ``anchor`` is any trainable parameter: it makes autograd call
    ``backward``; the real gradients are written into ``p.grad`` by the engine (see the module docstring)."""

    @staticmethod
    def forward(ctx, anchor, engine, ent, logits):
        ctx.engine, ctx.ent = engine, ent
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.engine.backward(ctx.ent, dlogits)
        return None, None, None, None


def run_train_step_forward(engine: TrainEngine, anchor: torch.Tensor, *args, **kwargs):
    """Forward through the engine, wrapped in the autograd node.  Returns (logits, scores, idxs)."""
    with torch.no_grad():
        logits, scores, idxs, ent = engine.forward(*args, **kwargs)
    out = _TrainFn.apply(anchor, engine, ent, logits)
    return out, scores, idxs


def drop_path_scales(rates: Sequence[float], B: int, device, timm_04_style: bool = False):
    """Per-block (attention branch, MLP branch) DropPath scale vectors [B] fp32, drawn in the reference's order with the
    reference's RNG calls so that a seeded generator gives the reference's masks: timm >= 0.9 ``DropPath``
    (``x.new_empty(shape).bernoulli_(keep_prob).div_(keep_prob)``, AudioMAE, models_vit.py:149) or timm 0.4.5's
    (``floor(keep_prob + rand(shape)) / keep_prob``, AST).  A rate of 0 draws nothing (drop_path returns early)."""
    out = []
    for r in rates:
        pair = []
        for _ in range(2):
            if r == 0.0:
                pair.append(None)
                continue
            keep_prob = 1.0 - r
            if timm_04_style:
                t = (keep_prob + torch.rand((B, 1, 1), dtype=torch.float32, device=device)).floor_().div_(keep_prob)
            else:
                t = torch.empty((B, 1, 1), dtype=torch.float32, device=device).bernoulli_(keep_prob)
                if keep_prob > 0.0:
                    t.div_(keep_prob)
            pair.append(t.reshape(B).contiguous())
        out.append(tuple(pair))
    return out


class GraphedTrainStep:
    """The whole fine-tune step -- forward, loss, backward (with its NCCL gradient buckets), FusedAdamW -- as ONE CUDA
    graph: ~440 kernel launches and all the Python in between become a single ``replay()``.

        step = GraphedTrainStep(model, optimizer, criterion, sample_x, sample_y, keep_rate_list=None)
        for x, y in loader:
            lr_sched.adjust_learning_rate(optimizer, ...)      # group["lr"] is re-read (pinned buffer) at every replay
            loss = step(x, y)                                  # device tensor, valid until the next call

    What is baked in at capture time: shapes, the pruning schedule (``keep_rate_list``; capture one instance per
    schedule, as the reference's per-epoch keep-rate schedule changes it), mask probabilities, DropPath RATES (the masks
    themselves are redrawn by the captured RNG kernels at every replay).  Requires ``FusedAdamW`` (its step counter lives
    on the device).  Single-process only for now: with torch.distributed initialised the constructor raises -- capturing
    the NCCL gradient buckets hung on this stack (torch 2.11 / NCCL 2.28.9, two B200s), so multi-GPU steps run eagerly."""

    def __init__(self, model, optimizer, criterion, sample_x: torch.Tensor, sample_y: torch.Tensor, keep_rate_list=None,
                 warmup: int = 3, **forward_kwargs):
        from .optim import FusedAdamW
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("GraphedTrainStep needs tpat.optim.FusedAdamW (device-side step counter, flat buffers)")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            raise RuntimeError("GraphedTrainStep: capturing the NCCL gradient all-reduce is not supported; run the step eagerly")
        self.model, self.opt, self.criterion = model, optimizer, criterion
        self.kw = dict(keep_rate_list=keep_rate_list, **forward_kwargs)
        self.x = sample_x.detach().clone()
        self.y = sample_y.detach().clone()
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on a side stream (torch's capture recipe)
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.replays = 0

    def _eager(self):
        loss = self.criterion(self.model(self.x, **self.kw), self.y)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor, copy: bool = True) -> torch.Tensor:
        """One training step on (x, y).  ``copy=False``: the caller has already written into ``self.x`` / ``self.y``."""
        if copy:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
        self.opt.sync_hyperparams()           # the captured H2D copy of the (lr, weight decay) table reads the pinned buffer now
        self.graph.replay()
        self.replays += 1
        return self.loss
