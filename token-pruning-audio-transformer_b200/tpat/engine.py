"""Host-side driver of ``tpat_forward``: weight packing, workspace, pruning schedule, CUDA graphs.

The model classes (models_vit.VisionTransformer, ast_models.ASTModel) keep their parameters as
ordinary ``nn.Parameter``s under the reference's state-dict names; this engine reads them,
keeps bf16 copies of the matrices for the tensor-core path, fills the C ``tpat_forward_args``
struct and makes ONE native call per forward.
"""
import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ForwardArgs, check, lib

# "bf16+score32": bf16 tensor-core kernels, but the pruning blocks compute their q / k projection and Q K^T of the score
# tiles in split-bf16 (three tcgen05.mma per product, ~fp32 accuracy): the token SELECTION follows the reference to its
# fp32 accuracy (given the same block input) for a few per cent of extra time.
PRECISIONS = ("bf16", "fp32", "bf16+score32")


def resolve_precision(p: Optional[str]) -> str:
    p = p or os.environ.get("TPAT_PRECISION", "bf16")
    if p not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {p!r}")
    return p


def pruning_schedule(n_patches: int, num_extra: int, keep_rates: Sequence[float], fuse_token: bool = False
                     ) -> Tuple[List[int], List[int]]:
    """(prune flags, k of each block's top-k / unchanged token count), following the reference exactly:
    ``num_left_tokens = math.ceil(keep_rate * (N - num_extra_tokens))`` in Python double arithmetic
    and top-k only ``if keep_rate < 1.0`` (models_vit.py:104-110; ast_models.py:116-121).  With ``fuse_token``
    (EViT, not in the reference forward) a block that drops tokens hands k + 1 tokens to the next one."""
    prune, keep, cur = [], [], n_patches
    for kr in keep_rates:
        N = cur + num_extra
        k = math.ceil(kr * (N - num_extra))
        assert k > 0, "num_left_tokens should be at least 1"      # models_vit.py:106
        if kr < 1.0:
            prune.append(1)
            keep.append(k)
            cur = k + (1 if (fuse_token and k < cur) else 0)
        else:
            prune.append(0)
            keep.append(cur)
    return prune, keep


def tokens_entering(n_patches: int, prune: Sequence[int], keep: Sequence[int], fuse_token: bool) -> List[int]:
    """Non-extra token count entering each block (= length of that block's score vector)."""
    out, cur = [], n_patches
    for p, k in zip(prune, keep):
        out.append(cur)
        cur = k + (1 if (p and fuse_token and k < cur) else 0) if p else cur
    return out


class EnginePool:
    """One ForwardEngine per CUDA device, created on first use under a lock.

    ``nn.DataParallel`` (the reference's AST training wrapper, ast/src/traintest.py:79,286) replicates a module by
    shallow-copying ``__dict__``: every replica therefore shares this pool object, and each replica thread asks it
    for the engine of ITS device -- packed weights, bf16 copies, workspace and graphs are never shared across GPUs."""

    def __init__(self, *engine_args):
        import threading
        self._args = engine_args
        self._lock = threading.Lock()
        self._engines: Dict[int, "ForwardEngine"] = {}
        self._train: Dict[int, object] = {}
        self.ln_fold = os.environ.get("TPAT_LN_FOLD", "0")
        self.graph_static_io = False

    def get(self, device: torch.device) -> "ForwardEngine":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with self._lock:
            e = self._engines.get(idx)
            if e is None:
                e = ForwardEngine(*self._args)
                e.ln_fold, e.graph_static_io = self.ln_fold, self.graph_static_io
                self._engines[idx] = e
            return e

    def get_train(self, device: torch.device):
        """The TrainEngine (fine-tune step: flat parameter / gradient buffers, saved activations) of ``device``."""
        from .train import TrainEngine
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with self._lock:
            e = self._train.get(idx)
            if e is None:
                e = self._train[idx] = TrainEngine(*self._args)
            return e

    def invalidate_inference(self) -> None:
        """Drop the inference engines' packed weights / bf16 copies / graphs (the weights changed without a version bump)."""
        with self._lock:
            for e in self._engines.values():
                e.invalidate()

    def invalidate(self) -> None:
        self.invalidate_inference()
        with self._lock:
            for e in self._train.values():
                e.mark_updated()


def param_key(module) -> tuple:
    """(data_ptr, _version) of every parameter tensor reachable from ``module`` -- also on an nn.DataParallel
    replica, whose ``parameters()`` is empty (its tensors are plain attributes listed in ``_former_parameters``)."""
    out = []
    for m in module.modules():
        params = m._parameters if m._parameters else getattr(m, "_former_parameters", {})   # replica: see replicate()
        for p in params.values():
            if p is not None:
                out.append((p.data_ptr(), p._version))
    return tuple(out)


class ForwardEngine:
    """One instance per model.  ``tensors`` maps a small fixed vocabulary of names to device tensors:
    patch_w [D,1,16,16], patch_b, extra_tok [extra,D], pos [extra+P,D], blocks[i].{ln1_g,...},
    norm_g/norm_b, head_ln_g/head_ln_b (AST), head_w, head_b."""

    def __init__(self, variant: int, depth: int, D: int, H: int, Dh: int):
        if depth > _lib.TPAT_MAX_DEPTH:
            raise ValueError(f"depth {depth} exceeds TPAT_MAX_DEPTH={_lib.TPAT_MAX_DEPTH}")
        self.variant, self.depth, self.D, self.H, self.Dh = variant, depth, D, H, Dh
        self.num_extra = 2 if variant == _lib.VARIANT_AST else 1
        self._pack_key = None
        self._fold: Dict[int, Dict[str, torch.Tensor]] = {}
        # LayerNorm fold (tpat_gemm_ln, bf16 path): "0" off (default), "qkv" norm1 -> qkv only, "all" norm1 and norm2.
        # Measured r01f (B = 64): the 20 LayerNorm launches it removes (0.39 ms) are paid back by the producer's extra
        # bf16 write + moments (proj +16 us, fc2 +13 us) and the consumer epilogues (qkv +10 us, fc1 +20 us: the GELU
        # epilogue has no slack) -- 5.01 vs 4.87 ms per step for "all", no measurable change for "qkv" -> off.
        self.ln_fold = os.environ.get("TPAT_LN_FOLD", "0")
        self._packed: Dict[str, object] = {}
        self._bf16: Dict[int, torch.Tensor] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._graphs: Dict[tuple, tuple] = {}
        self.graph_static_io = False     # see _run_graph
        self._score32 = False
        self.last_launch_count = 0
        self.last_pooled: Optional[torch.Tensor] = None

    # ---- weights -------------------------------------------------------------------------
    @staticmethod
    def _f32(t: torch.Tensor) -> torch.Tensor:
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        return t

    def pack(self, tensors_fn, key) -> None:
        """(Re)build the device-side weight views when ``key`` (data pointers + versions) changed.
        ``tensors_fn`` is only called on a miss."""
        if key == self._pack_key:
            return
        tensors = tensors_fn()
        f = self._f32
        pk: Dict[str, object] = {}
        pk["patch_w"] = f(tensors["patch_w"]).reshape(self.D, 256)
        pk["patch_b"] = f(tensors["patch_b"])
        pk["extra_tok"] = f(tensors["extra_tok"]).reshape(self.num_extra, self.D).contiguous()
        pk["pos"] = f(tensors["pos"]).reshape(-1, self.D)
        pk["blocks"] = [{k: f(v) for k, v in blk.items()} for blk in tensors["blocks"]]
        for k in ("norm_g", "norm_b", "head_ln_g", "head_ln_b", "head_w", "head_b"):
            pk[k] = f(tensors[k]) if tensors.get(k) is not None else None
        self._packed = pk
        self._fold = {}
        self._bf16 = {}
        self._graphs = {}
        self._pack_key = key

    def invalidate(self) -> None:
        """Drop the packed weight views, bf16 copies and captured graphs.  The pack key is (data_ptr, _version) per
        parameter; writes through ``p.data`` (EMA / weight averaging code) do not bump ``_version``, so such callers
        must call this (``model.invalidate_packed()``) after changing weights in place."""
        self._pack_key = None
        self._packed, self._fold, self._bf16, self._graphs = {}, {}, {}, {}

    def _mat(self, t: torch.Tensor, impl: int) -> torch.Tensor:
        if impl == _lib.IMPL_SIMT:
            return t
        c = self._bf16.get(id(t))
        if c is None:
            c = t.to(torch.bfloat16).contiguous()
            self._bf16[id(t)] = c
        return c

    def _qk_split(self, i: int) -> torch.Tensor:
        """[2D, 3D] bf16 = [w_hi | w_hi | w_lo] of the q and k rows of block i's qkv weight (pairs with the
        [hi | lo | hi] LayerNorm output: a_hi w_hi + a_lo w_hi + a_hi w_lo)."""
        key = ("qk_split", i)
        c = self._bf16.get(key)
        if c is None:
            w = self._packed["blocks"][i]["qkv_w"][: 2 * self.D]
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            c = torch.cat([hi, hi, lo], dim=1).contiguous()
            self._bf16[key] = c
        return c

    def _folded(self, i: int) -> Dict[str, torch.Tensor]:
        """Weights of the LayerNorm fold for block i (tpat_gemm_ln): W' = bf16(W * gamma), colsum = sum_k W', b' = W beta + b."""
        f = self._fold.get(i)
        if f is None:
            blk = self._packed["blocks"][i]
            f = {}
            for name, g, b in (("qkv", "ln1_g", "ln1_b"), ("fc1", "ln2_g", "ln2_b")):
                w = blk[name + "_w"]
                wl = (w * blk[g][None, :]).to(torch.bfloat16).contiguous()
                f[name + "_w_ln"] = wl
                f[name + "_colsum"] = wl.float().sum(dim=1).contiguous()
                f[name + "_b_ln"] = (w @ blk[b] + blk[name + "_b"]).contiguous()
            self._fold[i] = f
        return f

    # ---- one forward ---------------------------------------------------------------------
    def _fill_args(self, spec, prune, keep, want_all_scores, impl, num_classes, logits, scores, idxs, workspace,
                   fuse_token=False):
        pk = self._packed
        a = ForwardArgs()
        a.score32 = 1 if self._score32 else 0
        a.fuse_token = 1 if fuse_token else 0
        a.variant, a.impl = self.variant, impl
        a.B, a.T, a.F = spec.shape
        a.depth, a.D, a.H, a.Dh, a.num_classes = self.depth, self.D, self.H, self.Dh, num_classes
        for i in range(self.depth):
            a.prune[i], a.keep[i] = prune[i], keep[i]
        a.want_all_scores = 1 if want_all_scores else 0
        a.ln_eps = 1e-6
        a.patch_w = self._mat(pk["patch_w"], impl).data_ptr()
        a.patch_b = pk["patch_b"].data_ptr()
        a.extra_tok = pk["extra_tok"].data_ptr()
        a.pos = pk["pos"].data_ptr()
        for i, blk in enumerate(pk["blocks"]):
            bw = a.blocks[i]
            for name in ("ln1_g", "ln1_b", "qkv_b", "proj_b", "ln2_g", "ln2_b", "fc1_b", "fc2_b"):
                setattr(bw, name, blk[name].data_ptr())
            for name in ("qkv_w", "proj_w", "fc1_w", "fc2_w"):
                setattr(bw, name, self._mat(blk[name], impl).data_ptr())
            if impl == _lib.IMPL_TC and self.ln_fold in ("qkv", "all"):
                for name, t in self._folded(i).items():
                    if self.ln_fold == "all" or name.startswith("qkv"):
                        setattr(bw, name, t.data_ptr())
            if impl == _lib.IMPL_TC and self._score32 and prune[i]:
                bw.qk_w_split = self._qk_split(i).data_ptr()
        a.norm_g, a.norm_b, a.norm_eps = pk["norm_g"].data_ptr(), pk["norm_b"].data_ptr(), 1e-6
        if pk["head_ln_g"] is not None:
            a.head_ln_g, a.head_ln_b, a.head_ln_eps = pk["head_ln_g"].data_ptr(), pk["head_ln_b"].data_ptr(), 1e-5
        a.head_w, a.head_b = pk["head_w"].data_ptr(), pk["head_b"].data_ptr()
        a.spec, a.logits = spec.data_ptr(), logits.data_ptr()
        for i in range(self.depth):
            a.scores[i] = scores[i].data_ptr() if scores[i] is not None else None
            a.topk_idx[i] = idxs[i].data_ptr() if idxs[i] is not None else None
        if workspace is not None:
            a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
        return a

    def _alloc_outputs(self, spec, prune, keep, want_all_scores, num_classes, fuse_token=False):
        B, T, F = spec.shape
        dev = spec.device
        logits = torch.empty(B, num_classes, device=dev, dtype=torch.float32)
        scores: List[Optional[torch.Tensor]] = [None] * self.depth
        idxs: List[Optional[torch.Tensor]] = [None] * self.depth
        entering = tokens_entering((T // 16) * (F // 16), prune, keep, fuse_token)
        for i in range(self.depth):
            if prune[i] or want_all_scores:
                scores[i] = torch.empty(B, entering[i], device=dev, dtype=torch.float32)
            if prune[i]:
                idxs[i] = torch.empty(B, keep[i], device=dev, dtype=torch.int64)
        return logits, scores, idxs

    def _ensure_workspace(self, args: ForwardArgs, device) -> torch.Tensor:
        need = lib.tpat_forward_workspace_bytes(ctypes.byref(args))
        if need == 0:
            raise RuntimeError(f"libtpat tpat_forward_workspace_bytes failed: {_lib.last_error()}")
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = torch.empty(need, device=device, dtype=torch.uint8)
            self._graphs = {}
        return self._workspace

    def run(self, spec: torch.Tensor, *args, **kwargs):
        """spec [B,T,F] fp32 CUDA.  Returns (logits [B,C], scores list, topk_idx list) -- device tensors.
        Runs on the INPUT's device and that device's current stream (as a torch op would), so one process can drive
        several GPUs (the reference's nn.DataParallel replicas)."""
        if not spec.is_cuda:
            raise RuntimeError("tpat: input must be a CUDA tensor; there is no CPU path")
        with torch.cuda.device(spec.device):
            return self._run(spec, *args, **kwargs)

    def _run(self, spec: torch.Tensor, keep_rates: Sequence[float], num_classes: int, want_all_scores: bool = False,
             precision: str = "bf16", use_graph: bool = False, fuse_token: bool = False):
        if not lib.tpat_device_ok():
            raise RuntimeError("tpat: the current device is not compute capability 10.x (B200, sm_100a)")
        if spec.dtype != torch.float32 or not spec.is_contiguous():
            spec = spec.float().contiguous()
        impl = _lib.IMPL_SIMT if precision == "fp32" else _lib.IMPL_TC
        self._score32 = precision == "bf16+score32"
        B, T, F = spec.shape
        prune, keep = pruning_schedule((T // 16) * (F // 16), self.num_extra, keep_rates, fuse_token)
        if use_graph:
            return self._run_graph(spec, prune, keep, want_all_scores, impl, num_classes, fuse_token)
        logits, scores, idxs = self._alloc_outputs(spec, prune, keep, want_all_scores, num_classes, fuse_token)
        args = self._fill_args(spec, prune, keep, want_all_scores, impl, num_classes, logits, scores, idxs, None, fuse_token)
        ws = self._ensure_workspace(args, spec.device)
        args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
        pooled = torch.empty(B, self.D, device=spec.device, dtype=torch.float32)
        args.pooled = pooled.data_ptr()
        self.last_launch_count = lib.tpat_forward_launch_count(ctypes.byref(args))
        check(lib.tpat_forward(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "tpat_forward")
        self.last_pooled = pooled                            # the classifier input (forward_features' return value)
        return logits, scores, idxs

    def _run_graph(self, spec, prune, keep, want_all_scores, impl, num_classes, fuse_token=False):
        """CUDA-graph replay of ``tpat_forward``.  Two flavours (``self.graph_static_io``):

        * False (default): the graph reads a private static input; every call copies ``spec`` into it and returns
          clones of the outputs (safe for any caller, ~10 ATen launches per step).
        * True: the graph reads the CALLER's tensor in place (one graph per distinct input buffer, keyed by its
          address) and the outputs are returned as views of the graph's own output buffers, valid until the next
          replay of the same graph -- a step is ``g.replay()`` and nothing else (no ATen kernel on the path).  This is
          what a serving loop with fixed staging buffers (bench.py) uses."""
        static_io = bool(self.graph_static_io)
        key = (tuple(spec.shape), tuple(prune), tuple(keep), bool(want_all_scores), impl, num_classes, spec.device.index,
               bool(fuse_token), spec.data_ptr() if static_io else None, self._score32)
        ent = self._graphs.get(key)
        if ent is None:
            if static_io:
                static_in = spec
            else:
                static_in = torch.empty_like(spec)
                static_in.copy_(spec)
            logits, scores, idxs = self._alloc_outputs(spec, prune, keep, want_all_scores, num_classes, fuse_token)
            args = self._fill_args(static_in, prune, keep, want_all_scores, impl, num_classes, logits, scores, idxs, None,
                                   fuse_token)
            ws = self._ensure_workspace(args, spec.device)
            args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
            pooled = torch.empty(spec.shape[0], self.D, device=spec.device, dtype=torch.float32)
            args.pooled = pooled.data_ptr()
            self.last_launch_count = lib.tpat_forward_launch_count(ctypes.byref(args))
            # warm-up outside capture (function attributes, tensor-map cache), then capture
            check(lib.tpat_forward(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "tpat_forward")
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                check(lib.tpat_forward(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "tpat_forward")
            ent = (g, static_in, logits, scores, idxs, ws, pooled)
            if len(self._graphs) >= 16:                      # bounded cache: drop the oldest captured schedule
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = ent
        g, static_in, logits, scores, idxs, _, pooled = ent
        if static_io:
            g.replay()
            self.last_pooled = pooled
            return logits, list(scores), list(idxs)
        static_in.copy_(spec)
        g.replay()
        self.last_pooled = pooled.clone()
        return (logits.clone(), [None if s is None else s.clone() for s in scores],
                [None if t is None else t.clone() for t in idxs])

    # ---- kernel-by-kernel forward: ablation and masking paths -------------------------------
    def run_stepwise(self, spec: torch.Tensor, *args, **kwargs):
        """Device-guarded entry of ``_run_stepwise`` (see there)."""
        if not spec.is_cuda:
            raise RuntimeError("tpat: input must be a CUDA tensor; there is no CPU path")
        with torch.cuda.device(spec.device):
            return self._run_stepwise(spec, *args, **kwargs)

    def _run_stepwise(self, spec: torch.Tensor, keep_rates: Sequence[float], num_classes: int, precision: str = "bf16",
                      use_custom_rank: Optional[str] = None, drop_token_blk_idx: Optional[int] = None,
                      retain_min: Optional[float] = None, retain_max: Optional[float] = None,
                      mask_keep_idx: Optional[torch.Tensor] = None):
        """The forward assembled from the per-kernel entry points, for the paths whose token flow the fused
        ``tpat_forward`` does not cover:

        * ``use_custom_rank`` 'mean' | 'std': rank tokens by a statistic of their spectrogram patch; the gather takes
          its indices over the FULL token list and keeps exactly k tokens, as the reference's
          ``forward_with_custom_rank`` does (models_vit.py:209-224,343-351,371-374; ast_models.py:221-236,445-453).
        * ``drop_token_blk_idx``: batch of one; after that block keep the patches whose mean intensity lies in
          (retain_min, retain_max); returns ``None`` when none is left (models_vit.py:353-355,378-385).
        * ``mask_keep_idx`` [B, n_keep] int64: patch tokens that survive the fine-tune 2-D masking, applied right
          after the patch embedding (models_vit.py:425-497).

        Same kernels as ``run`` (every op in libtpat.so); only the sequencing lives here.  Returns
        (logits or None, {'topk_idx': {block: idx}, 'retain_idx': idx or None})."""
        from . import ops
        if not spec.is_cuda:
            raise RuntimeError("tpat: input must be a CUDA tensor; there is no CPU path")
        if not lib.tpat_device_ok():
            raise RuntimeError("tpat: the current device is not compute capability 10.x (B200, sm_100a)")
        if use_custom_rank not in (None, "mean", "std"):
            raise ValueError(f"custom_rank should be in ['mean', 'std'], got {use_custom_rank}")
        if spec.dtype != torch.float32 or not spec.is_contiguous():
            spec = spec.float().contiguous()
        impl = _lib.IMPL_SIMT if precision == "fp32" else _lib.IMPL_TC      # (the stepwise paths have no score32 variant)
        act = torch.bfloat16 if impl == _lib.IMPL_TC else torch.float32
        pk, extra, D, H = self._packed, self.num_extra, self.D, self.H
        ast = self.variant == _lib.VARIANT_AST
        order = _lib.TOKENS_FREQ_MAJOR if ast else _lib.TOKENS_TIME_MAJOR
        B, T, F = spec.shape
        P = (T // 16) * (F // 16)
        info = {"topk_idx": {}, "retain_idx": None}

        rank = None
        if use_custom_rank is not None:
            mean, std = ops.patch_stats(spec, order, want_mean=use_custom_rank == "mean", want_std=use_custom_rank == "std")
            rank = mean if use_custom_rank == "mean" else std
        intensity = ops.patch_stats(spec, order)[0] if drop_token_blk_idx is not None else None

        x = torch.empty(B, extra + P, D, device=spec.device, dtype=torch.float32)
        patches = ops.patchify(spec, act, order, tokens=x, extra_tok=pk["extra_tok"], pos=pk["pos"])
        ops.gemm(patches, self._mat(pk["patch_w"], impl), pk["patch_b"], torch.float32, _lib.EPI_BIAS_POS, impl,
                 out=x.view(-1, D), pos=pk["pos"], P=P, num_extra=extra)
        if mask_keep_idx is not None:
            x, _ = ops.gather_layernorm(x, mask_keep_idx.contiguous(), extra, None, None, 1e-6, act)

        for i, blk in enumerate(pk["blocks"]):
            N = x.shape[1]
            kr = keep_rates[i]
            k = math.ceil(kr * (N - extra))
            assert k > 0                                                         # models_vit.py:106
            prune = kr < 1.0
            y = ops.layernorm(x, blk["ln1_g"], blk["ln1_b"], 1e-6, act)
            qkv = ops.gemm(y.view(-1, D), self._mat(blk["qkv_w"], impl), blk["qkv_b"], act, _lib.EPI_BIAS, impl)
            smode = _lib.SCORE_NONE
            if prune and rank is None:
                smode = _lib.SCORE_CLS_ROW if ast else _lib.SCORE_COLMEAN
            ao, partial = ops.attention(qkv, B, N, H, extra, smode, impl)
            x2 = x.view(-1, D)
            ops.gemm(ao, self._mat(blk["proj_w"], impl), blk["proj_b"], torch.float32, _lib.EPI_BIAS_RESIDUAL, impl,
                     residual=x2, out=x2)
            if prune and rank is not None:
                _, idx = ops.score_topk(rank.view(B, 1, -1), 1.0, 0, k)          # torch.topk(custom_rank, k)
                x, y2 = ops.gather_layernorm(x, idx, 0, blk["ln2_g"], blk["ln2_b"], 1e-6, act)   # full token list
                rank = ops.gather_rank(rank, idx)
                info["topk_idx"][i] = idx
            elif prune:
                divisor = float(H) if ast else float(H) * float(N - extra)
                _, idx = ops.score_topk(partial, divisor, extra, k)
                x, y2 = ops.gather_layernorm(x, idx, extra, blk["ln2_g"], blk["ln2_b"], 1e-6, act)
                info["topk_idx"][i] = idx
            else:
                y2 = ops.layernorm(x, blk["ln2_g"], blk["ln2_b"], 1e-6, act)
            hdn = ops.gemm(y2.view(-1, D), self._mat(blk["fc1_w"], impl), blk["fc1_b"], act, _lib.EPI_BIAS_GELU, impl)
            x2 = x.view(-1, D)
            ops.gemm(hdn, self._mat(blk["fc2_w"], impl), blk["fc2_b"], torch.float32, _lib.EPI_BIAS_RESIDUAL, impl,
                     residual=x2, out=x2)
            if drop_token_blk_idx == i:
                assert B == 1                                                    # models_vit.py:379
                m = intensity[0].cpu()                                           # host decision: data-dependent token count
                retain = torch.nonzero((m > retain_min) & (m < retain_max))[:, 0]
                if retain.numel() == 0:
                    return None, info
                if int(retain.max()) + extra >= x.shape[1]:
                    raise IndexError(f"index {int(retain.max()) + extra} is out of bounds for dimension 1 with size "
                                     f"{x.shape[1]} (the intensity filter indexes the original patch grid)")
                info["retain_idx"] = retain
                x, _ = ops.gather_layernorm(x, retain.to(x.device).view(1, -1).contiguous(), extra, None, None, 1e-6, act)

        pooled = ops.pool_norm(x, self.variant, pk["norm_g"], pk["norm_b"], 1e-6, pk["head_ln_g"], pk["head_ln_b"], 1e-5)
        return ops.head(pooled, pk["head_w"], pk["head_b"]), info
