"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/*.pt from the REAL reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For every entry of ``golden_configs.GOLDEN_CONFIGS`` it loads the deterministic state-dict
into the unmodified reference model (oracle/ref_loader.py), runs the reference forward in
extract mode (fp32, eval, no_grad) and stores

    ref      logits + 'block-i.attn_score' (12 blocks) + 'block-i.topk_idx' from the reference
    ref_plain  logits from the non-extract call ``model(x)``
    f64      the same quantities from the oracle restatement in float64 (tie ranking / "exact")
    meta     the config, torch version and the sha256 digest of the state-dict and the input

The reference has no tests or golden vectors of its own, so these files
are the parity pin.
"""
import hashlib
import os
import sys

import torch

from . import ref_loader, weights, vit_oracle as vo
from .golden_configs import ABLATION_CONFIGS, GOLDEN_CONFIGS, GRAD_CONFIGS, MASKED_CONFIGS

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_inputs(cfg):
    if cfg["variant"] == "audiomae":
        sd = weights.make_audiomae_state_dict(cfg["num_classes"], cfg["T"], cfg["wseed"], cfg["flavour"])
    else:
        sd = weights.make_ast_state_dict(cfg["num_classes"], cfg["T"], cfg["wseed"], cfg["flavour"])
    x = weights.make_spectrogram(cfg["variant"], cfg["B"], cfg["T"], cfg["xseed"])
    return sd, x


def build_reference(cfg):
    if cfg["variant"] == "audiomae":
        return ref_loader.build_audiomae(cfg["num_classes"], cfg["T"], cfg["drop_loc"], cfg["base_keep_rate"])
    return ref_loader.build_ast(cfg["num_classes"], cfg["T"], cfg["drop_loc"], cfg["base_keep_rate"])


def main(names=None):
    assert ref_loader.reference_available(), "run in the build container (needs /root/reference)"
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, cfg in GOLDEN_CONFIGS.items():
        if names and name not in names:
            continue
        sd, x = make_inputs(cfg)
        model = build_reference(cfg)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        krl = cfg["keep_rate_list"]
        with torch.no_grad():
            logits, feats = model(x, keep_rate_list=krl, flag_extract_features=True)
            logits_plain = model(x, keep_rate_list=krl)
            o_logits, o_feats = vo.forward(cfg["variant"], sd, x, krl, cfg["drop_loc"], cfg["base_keep_rate"],
                                           flag_extract_features=True)
            d_logits, d_feats = vo.forward(cfg["variant"], sd, x, krl, cfg["drop_loc"], cfg["base_keep_rate"],
                                           flag_extract_features=True, dtype=torch.float64)
        feats = {k: v for k, v in feats.items() if k != "mel"}
        # the restatement must be bit-identical to the reference before anything is written
        assert torch.equal(logits, o_logits) and torch.equal(logits, logits_plain), name
        assert sorted(feats) == sorted(o_feats), name
        for k in feats:
            assert torch.equal(feats[k], o_feats[k]), (name, k)
        blob = {
            "meta": dict(cfg, name=name, torch=torch.__version__,
                         sd_digest=weights.state_dict_digest(sd),
                         x_digest=hashlib.sha256(x.numpy().tobytes()).hexdigest()),
            "ref": {"logits": logits, **feats},
            "ref_plain": {"logits": logits_plain},
            "f64": {"logits": d_logits, **d_feats},
        }
        path = os.path.join(OUT_DIR, name + ".pt")
        torch.save(blob, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), logits {tuple(logits.shape)}")


def main_ablation(names=None):
    """tests/golden/abl_<name>.pt: the reference's output on its ablation paths."""
    assert ref_loader.reference_available(), "run in the build container (needs /root/reference)"
    for name, cfg in ABLATION_CONFIGS.items():
        if names and name not in names:
            continue
        sd, x = make_inputs(cfg)
        model = build_reference(cfg)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        model.use_custom_rank = cfg.get("use_custom_rank")                 # main_finetune.py:448-455 / run.py:204-211
        if cfg.get("drop_token_blk_idx") is not None:
            model.retain_min, model.retain_max = cfg["retain_min"], cfg["retain_max"]
            model.drop_token_blk_idx = cfg["drop_token_blk_idx"]
        kw = dict(use_custom_rank=cfg.get("use_custom_rank"), drop_token_blk_idx=cfg.get("drop_token_blk_idx"),
                  retain_min=cfg.get("retain_min"), retain_max=cfg.get("retain_max"))
        with torch.no_grad():
            logits = model(x, keep_rate_list=cfg["keep_rate_list"])
            o_logits, info = vo.forward_ablation(cfg["variant"], sd, x, cfg["keep_rate_list"], cfg["drop_loc"],
                                                 cfg["base_keep_rate"], **kw)
            d_logits, d_info = vo.forward_ablation(cfg["variant"], sd, x, cfg["keep_rate_list"], cfg["drop_loc"],
                                                   cfg["base_keep_rate"], dtype=torch.float64, **kw)
        assert (logits is None) == (o_logits is None), name
        if logits is not None:
            assert torch.equal(logits, o_logits), name            # restatement bit-identical to the reference
        blob = {
            "meta": dict(cfg, name=name, torch=torch.__version__, sd_digest=weights.state_dict_digest(sd),
                         x_digest=hashlib.sha256(x.numpy().tobytes()).hexdigest()),
            "ref": {"logits": logits},
            "oracle_info": info,       # decisions of the fp32 restatement (the reference does not return them)
            "f64": {"logits": d_logits, "info": d_info},
        }
        path = os.path.join(OUT_DIR, "abl_" + name + ".pt")
        torch.save(blob, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), logits "
              f"{None if logits is None else tuple(logits.shape)}")


def main_masked(names=None):
    """tests/golden/mask_<name>.pt: reference logits of forward(x, mask_t_prob, mask_f_prob) in eval mode."""
    assert ref_loader.reference_available(), "run in the build container (needs /root/reference)"
    for name, cfg in MASKED_CONFIGS.items():
        if names and name not in names:
            continue
        sd, x = make_inputs(cfg)
        model = build_reference(cfg)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        B, Tp = cfg["B"], cfg["T"] // 16
        torch.manual_seed(cfg["mseed"])
        noise_t, noise_f = torch.rand(B, Tp), torch.rand(B, 8)               # the draws random_masking_2d will make
        torch.manual_seed(cfg["mseed"])
        with torch.no_grad():
            logits = model(x, keep_rate_list=cfg["keep_rate_list"], mask_t_prob=cfg["mask_t_prob"], mask_f_prob=cfg["mask_f_prob"])
            keep_idx = vo.masking_2d_keep_indices(noise_t, noise_f, cfg["mask_t_prob"], cfg["mask_f_prob"])
            o_logits = vo.forward_masked(cfg["variant"], sd, x, keep_idx, cfg["keep_rate_list"], cfg["drop_loc"],
                                         cfg["base_keep_rate"])
        assert torch.equal(logits, o_logits), name                            # restatement bit-identical to the reference
        blob = {"meta": dict(cfg, name=name, torch=torch.__version__, sd_digest=weights.state_dict_digest(sd),
                             x_digest=hashlib.sha256(x.numpy().tobytes()).hexdigest()),
                "noise_t": noise_t, "noise_f": noise_f, "keep_idx": keep_idx, "ref": {"logits": logits}}
        path = os.path.join(OUT_DIR, "mask_" + name + ".pt")
        torch.save(blob, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), kept {keep_idx.shape[1]} of {Tp * 8} patches")


def grad_summary(g: torch.Tensor) -> dict:
    """Compact fingerprint of one gradient tensor: l2 norm, sum, and 64 samples at a fixed stride."""
    f = g.detach().reshape(-1).double()
    stride = max(1, f.numel() // 64)
    return {"norm": f.norm().item(), "sum": f.sum().item(), "numel": f.numel(), "samples": f[::stride][:64].clone()}


def make_targets(B, C, seed):
    return (torch.rand(B, C, generator=torch.Generator().manual_seed(seed)) < 0.1).float()


def train_rng_draws(cfg, rates):
    """The random draws of one reference training forward, in its order: 2-D masking noise (models_vit.py:425-463:
    time first, then frequency), then the DropPath masks block by block."""
    torch.manual_seed(cfg["dseed"])
    keep_idx = None
    if cfg["mask_t_prob"] > 0 or cfg["mask_f_prob"] > 0:
        noise_t, noise_f = torch.rand(cfg["B"], cfg["T"] // 16), torch.rand(cfg["B"], 8)
        keep_idx = vo.masking_2d_keep_indices(noise_t, noise_f, cfg["mask_t_prob"], cfg["mask_f_prob"])
    return keep_idx, vo.drop_path_scales(rates, cfg["B"])


def main_grads(names=None):
    """tests/golden/grad_<name>.pt: gradients of the real reference's fine-tune step (train mode, autograd)."""
    assert ref_loader.reference_available(), "run in the build container (needs /root/reference)"
    for name, cfg in GRAD_CONFIGS.items():
        if names and name not in names:
            continue
        sd, x = make_inputs(cfg)
        y = make_targets(cfg["B"], cfg["num_classes"], cfg["tseed"])
        model = build_reference(cfg)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        model.train()
        blocks = model.blocks if cfg["variant"] == "audiomae" else model.v.blocks
        rates = [float(getattr(b.drop_path, "drop_prob", 0.0)) for b in blocks]
        torch.manual_seed(cfg["dseed"])
        if cfg["variant"] == "audiomae":
            logits = model(x, keep_rate_list=cfg["keep_rate_list"], mask_t_prob=cfg["mask_t_prob"], mask_f_prob=cfg["mask_f_prob"])
        else:
            logits = model(x, keep_rate_list=cfg["keep_rate_list"])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
        loss.backward()
        ref_grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
        # the restatement under autograd must reproduce them (same draws, fp32): pin before writing
        keep_idx, scales = train_rng_draws(cfg, rates)
        frozen = ("pos_embed",) if cfg["variant"] == "audiomae" else ()
        o_loss, o_logits, o_grads = vo.loss_and_grads(cfg["variant"], sd, x, y, cfg["keep_rate_list"], cfg["drop_loc"],
                                                      cfg["base_keep_rate"], dtype=torch.float32, drop_scales=scales,
                                                      mask_keep_idx=keep_idx, frozen=frozen)
        assert torch.allclose(logits.detach(), o_logits, rtol=0, atol=2e-6 * logits.abs().max().item()), name
        assert sorted(ref_grads) == sorted(o_grads), (name, sorted(set(ref_grads) ^ set(o_grads)))
        worst = max(((ref_grads[k] - o_grads[k]).norm() / ref_grads[k].norm().clamp_min(1e-30)).item() for k in ref_grads)
        assert worst < 1e-4, (name, worst)
        blob = {"meta": dict(cfg, name=name, torch=torch.__version__, sd_digest=weights.state_dict_digest(sd),
                             x_digest=hashlib.sha256(x.numpy().tobytes()).hexdigest(), drop_rates=rates),
                "loss": loss.item(), "logits": logits.detach(), "keep_idx": keep_idx,
                "grads": {k: grad_summary(v) for k, v in ref_grads.items()}}
        path = os.path.join(OUT_DIR, "grad_" + name + ".pt")
        torch.save(blob, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), {len(ref_grads)} gradient tensors, loss {loss.item():.6f}, "
              f"oracle(fp32 autograd) vs reference worst rel diff {worst:.2e}")


def main_schedule():
    """tests/golden/keep_rate_schedule.pt: outputs of the reference's OWN ``get_scheduled_keep_rate_list``
    (audiomae/engine_finetune.py:29-53).  The module cannot be imported as a whole here (it needs timm.data.Mixup),
    so the one function is cut out of the unmodified source with ``ast`` and compiled on its own."""
    import ast
    import math
    src_path = os.path.join(ref_loader.REFERENCE_ROOT, "audiomae", "engine_finetune.py")
    tree = ast.parse(open(src_path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_scheduled_keep_rate_list")
    ns = {"math": math}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, "exec"), ns)
    ref_fn = ns["get_scheduled_keep_rate_list"]
    cases = []
    for epoch in (0, 4, 5, 7, 12, 14, 15, 20):
        for off in (0, 11, 36):
            kw = dict(iters=epoch * 37 + off, epoch=epoch, shrink_start_epoch=5, total_epochs=15, ITERS_PER_EPOCH=37,
                      base_keep_rate=0.7, num_blocks=12, drop_loc=(3, 6, 9))
            cases.append((kw, ref_fn(**kw)))
    path = os.path.join(OUT_DIR, "keep_rate_schedule.pt")
    torch.save({"cases": cases}, path)
    print(f"keep_rate_schedule: wrote {path} ({len(cases)} cases)")


def main_fbank():
    """tests/golden/fbank_cases.pt: the eval-time input pipeline (audiomae/dataset.py:175-178,209-229,298) run through
    the reference's pinned third-party dependency, torchaudio.compliance.kaldi.fbank, on three deterministic waveforms
    (shorter than / exactly / longer than the target length).  The dataset class itself reads audio files and json
    manifests and cannot be driven without them; oracle/fbank_oracle.py restates the three statements around the
    torchaudio call."""
    import torchaudio
    from . import fbank_oracle as fo
    cases = {"short_1s": (16000, 1, 128), "exact_2p06s": (32960, 2, 204), "long_1p5s_crop": (24000, 3, 100)}
    blob = {"meta": {"torchaudio": torchaudio.__version__, "torch": torch.__version__, "cases": list(cases)}}
    for name, (n, seed, T) in cases.items():
        x = fo.make_waveform(n, seed)
        blob[name] = {"n": n, "seed": seed, "T": T, "x_digest": hashlib.sha256(x.numpy().tobytes()).hexdigest(),
                      "spec": fo.wav2fbank(x, target_length=T)}
    path = os.path.join(OUT_DIR, "fbank_cases.pt")
    torch.save(blob, path)
    print(f"fbank_cases: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "grads":
        main_grads(args[1:])
        sys.exit(0)
    if args and args[0] == "schedule":
        main_schedule()
        sys.exit(0)
    if args and args[0] == "fbank":
        main_fbank()
        sys.exit(0)
    if args and args[0] == "masked":
        main_masked(args[1:])
        sys.exit(0)
    if args and args[0] == "ablation":
        main_ablation(args[1:])
    else:
        main(args)
