"""FusedAdamW: torch.optim.AdamW's update as ONE kernel launch over the TrainEngine's flat buffers.

The reference's optimizer is ``torch.optim.AdamW(param_groups_lrd(...), lr, betas=(0.9, 0.95))``
(audiomae/main_finetune.py:464-468): ~28 parameter groups (layer id x decay / no-decay) over 151 tensors.  Here the
parameters, gradients and both moments are flat fp32 buffers; a chunk table maps every 16 Ki-element chunk to its
group, so one ``tpat_adamw`` launch updates everything, folds the 1 / world gradient averaging in, and refreshes the
bf16 operand copy of the weights that the next forward's tensor-core GEMMs read.  It is a ``torch.optim.Optimizer``
(``param_groups`` with the usual keys), so learning-rate schedulers and ``GradScaler.step`` work unchanged."""
from typing import List

import torch

from . import _lib
from ._lib import check, lib

CHUNK = 16384


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, model=None):
        if model is None:
            raise ValueError("FusedAdamW needs model= (the tpat model whose TrainEngine owns the flat buffers)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self._engine = None
        self._step = 0
        self._step_dev = None            # device-side step counter (CUDA-graph replays of the whole step)
        self._m = self._v = None
        self._chunks = self._groups_dev = self._groups_host = None
        self._n_chunks = 0

    def _bind(self):
        engine = self.model._engines.get_train(self.model._device_of_params())
        if engine.flat_p is None:      # no training forward has run yet: flatten now
            roles = self.model._train_roles()
            engine.attach(self.model._train_entries(roles), {})
        if engine is self._engine and self._m is not None and self._m.numel() == engine.flat_p.numel() \
                and self._flat_ptr == engine.flat_p.data_ptr():
            return engine
        dev = engine.flat_p.device
        rows: List[List[int]] = []
        covered = 0
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if id(p) not in engine.offsets:
                    if p.requires_grad and p.grad is not None:
                        raise RuntimeError("FusedAdamW: a parameter with a gradient is not part of the flat buffers")
                    continue                   # parameters the forward does not use (AST's v.head / v.head_dist)
                off, n = engine.offsets[id(p)]
                covered += 1
                for c0 in range(0, n, CHUNK):
                    rows.append([off + c0, min(CHUNK, n - c0), gi, 0])
        if covered != len(engine.offsets):
            raise RuntimeError("FusedAdamW: the parameter groups do not cover every trainable parameter of the model")
        self._chunks = torch.tensor(rows, dtype=torch.int32).to(dev)
        self._n_chunks = len(rows)
        self._groups_host = torch.zeros(len(self.param_groups), 2, dtype=torch.float32).pin_memory()
        self._groups_dev = torch.zeros(len(self.param_groups), 2, dtype=torch.float32, device=dev)
        if self._m is None or self._m.numel() != engine.flat_p.numel():
            self._m = torch.zeros_like(engine.flat_p)
            self._v = torch.zeros_like(engine.flat_p)
            self._step = 0
        self._engine, self._flat_ptr = engine, engine.flat_p.data_ptr()
        self._step_dev = torch.full((1,), self._step, dtype=torch.int32, device=dev)
        engine.fused_grad_scale = True
        return engine

    def sync_hyperparams(self) -> None:
        """param_groups -> the pinned (lr, weight decay) table the kernel's group table is copied from.  A CUDA graph of the
        step replays that copy, so a scheduler's new ``group["lr"]`` takes effect if this is called before the replay."""
        if self._groups_host is None:
            self._bind()
        for gi, g in enumerate(self.param_groups):                 # absolute per-group lr (what lr schedulers write)
            self._groups_host[gi, 0] = float(g["lr"])
            self._groups_host[gi, 1] = float(g["weight_decay"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        engine = self._bind()
        self._step += 1
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        self.sync_hyperparams()
        self._groups_dev.copy_(self._groups_host, non_blocking=True)
        pb = engine.flat_pb.data_ptr() if engine.flat_pb is not None else None
        stream = torch.cuda.current_stream().cuda_stream
        # the step count of the bias corrections lives on the device (one-thread kernel), so that a captured step replays
        check(lib.tpat_counter_inc(self._step_dev.data_ptr(), stream), "tpat_counter_inc")
        check(lib.tpat_adamw(engine.flat_p.data_ptr(), engine.flat_g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), pb,
                             self._chunks.data_ptr(), self._n_chunks, self._groups_dev.data_ptr(), 1.0, float(beta1), float(beta2),
                             float(eps), 0, self._step_dev.data_ptr(), float(engine.pending_grad_scale), stream), "tpat_adamw")
        engine.pending_grad_scale = 1.0
        engine.mark_updated(bf16_fresh=pb is not None)
        self.model._engines.invalidate_inference()     # the kernel wrote the weights without bumping tensor versions
        return loss

    def zero_grad(self, set_to_none: bool = True):
        """One memset of the flat gradient buffer (the ``p.grad`` views stay attached)."""
        engine = self._engine or self.model._engines.get_train(self.model._device_of_params())
        if engine.flat_g is not None:
            engine.flat_g.zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    def state_dict(self):
        step = int(self._step_dev.item()) if self._step_dev is not None else self._step
        return {"step": step, "m": self._m, "v": self._v,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._bind()
        self._step = int(sd["step"])
        self._step_dev.fill_(self._step)
        self._m.copy_(sd["m"]); self._v.copy_(sd["v"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
