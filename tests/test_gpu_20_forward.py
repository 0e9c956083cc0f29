"""GPU parity of the whole forward through the reference-facing model API against the golden
vectors produced by the REAL reference (tests/golden, oracle/make_golden.py) and the oracle.

fp32 mode  (CUDA-core kernels):  kept top-k SETS identical per clip and block, in mel-patch
           coordinates, except where the reference's own score gap at the cut is below 1e-9
           (fp32 summation order); logits within 1e-5 of max|logit| (north_star fp32 tolerance:
           2e-5 used).
bf16 mode  (tcgen05 kernels, bf16 operands / fp32 accumulate, fp32 residual + LN + softmax + score):
           SURVEY.md F14/H1 measured on the reference itself that random-init weights give almost
           uniform scores (cut gap ~4e-7), so its OWN bf16 run reaches only 0.97-0.99 kept-set
           overlap and 3.8e-2 logit error against its fp64 run; an emulated "bf16 operands, fp32
           accumulate" reference reaches 0.9983/0.9953/0.9929 and 2.0e-2.  The
           north-star 99.9 % / 1e-2 is therefore not reachable by ANY bf16 implementation on these
           inputs.  Asserted here (measured values are printed and recorded in DESIGN.md):
             * block-0 score within 5e-3 of max|score| (same tokens on both sides);
             * first pruning block: every token kept on one side only has a reference fp64 score
               within 1 % of the cut score (a near-tie flip, not a wrong selection);
             * every block: kept-set overlap (mel coordinates, vs the fp64 oracle) >= the value MEASURED for that golden
               config minus two flipped tokens (BF16_MEASURED below);
             * logits within 1.6x the measured error of that config, and within the north-star 1e-2 wherever that was
               measured (trained weight statistics, unpruned, ablation and masked paths).
           The per-block statement the kernels CAN be held to -- >= 99.9 % overlap given the reference's block input,
           with the "bf16+score32" split-precision score path -- is tests/test_gpu_25_parity_protocol.py.
"""
import pytest
import torch
import torch.nn as nn

import conftest  # noqa: F401
from conftest import load_golden, make_case
from gpu_util import dev, rel_err, set_overlap
from oracle import vit_oracle as vo
from oracle.golden_configs import GOLDEN_CONFIGS

pytestmark = pytest.mark.gpu


def build_model(meta, sd, precision):
    from tpat import models_vit, ASTModel
    if meta["variant"] == "audiomae":
        m = models_vit.vit_base_patch16(num_classes=meta["num_classes"], drop_path_rate=0.1, mean_pooling=True,
                                        mask_2d=True, target_length=meta["T"], drop_loc=tuple(meta["drop_loc"]),
                                        base_keep_rate=meta["base_keep_rate"], precision=precision)
        m.patch_embed = models_vit.PatchEmbed((meta["T"], 128), 16, 1, 768)
        m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, 768), requires_grad=False)
        m.load_state_dict(sd, strict=True)
    else:
        m = ASTModel(label_dim=meta["num_classes"], input_tdim=meta["T"], imagenet_pretrain=False,
                     audioset_pretrain=False, verbose=False, drop_loc=tuple(meta["drop_loc"]),
                     base_keep_rate=meta["base_keep_rate"], precision=precision)
        m.load_state_dict(sd, strict=False)
    return m.to(dev()).eval()


def prune_blocks(ref):
    return sorted(int(k.split(".")[0].split("-")[1]) for k in ref if k.endswith("topk_idx"))


def mel_idx(feats, blocks):
    return vo.melspec_indices([feats[f"block-{i}.topk_idx"].cpu() for i in blocks])


def check_fp32_sets(feats, ref, blocks):
    """Set equality per clip/block in mel coordinates, modulo near-ties at the cut in the reference."""
    if not blocks:
        return
    got, exp = mel_idx(feats, blocks), mel_idx(ref, blocks)
    for bi, (g, e) in enumerate(zip(got, exp)):
        blk = blocks[bi]
        score = ref[f"block-{blk}.attn_score"]
        k = e.shape[1]
        for c in range(e.shape[0]):
            sg, se = set(g[c].tolist()), set(e[c].tolist())
            if sg == se:
                continue
            srt = torch.sort(score[c], descending=True).values
            gap = (srt[k - 1] - srt[k]).item() if k < srt.numel() else float("inf")
            assert gap < 1e-9 and len(sg ^ se) <= 2, f"block {blk} clip {c}: kept sets differ, cut gap {gap:.3e}"
            return  # later blocks of this config are no longer comparable once a near-tie flipped


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_forward_fp32_matches_reference_golden(name):
    g = load_golden(name)
    meta, ref = g["meta"], g["ref"]
    sd, x = make_case(meta)
    model = build_model(meta, sd, "fp32")
    with torch.no_grad():
        logits, feats = model(x.to(dev()), keep_rate_list=meta["keep_rate_list"], flag_extract_features=True)
        logits_plain = model(x.to(dev()), keep_rate_list=meta["keep_rate_list"])
    assert sorted(k for k in feats if k != "mel") == sorted(k for k in ref if k != "logits")
    assert torch.equal(feats["mel"], x if meta["variant"] == "audiomae" else x.unsqueeze(1).transpose(2, 3))
    blocks = prune_blocks(ref)
    for i in blocks:
        t = feats[f"block-{i}.topk_idx"]
        assert t.dtype == torch.int64 and t.shape == ref[f"block-{i}.topk_idx"].shape and t.device.type == "cpu"
    check_fp32_sets(feats, ref, blocks)
    first = blocks[0] if blocks else 12
    for i in range(12):   # scores of blocks up to and including the first pruning block see identical tokens
        if i <= first:
            assert rel_err(feats[f"block-{i}.attn_score"], ref[f"block-{i}.attn_score"]) < 2e-5, i
    err = rel_err(logits.cpu(), ref["logits"])
    print(f"[fp32] {name}: logits err {err:.2e}")
    assert err < 2e-5
    assert torch.equal(logits, logits_plain)            # extract mode changes outputs reported, not math


# Measured on B200 (profiles/r01g_parity.txt, r02 parity file): per golden config, kept-set overlap vs the fp64 oracle at
# each pruning block and logits error / max|logit| of the plain bf16 mode.  The asserts below hold the kernels to
# these measurements (one or two more flipped tokens, 1.5x the logit error) instead of a blanket bound, so a
# regression from 0.995 to 0.97 overlap or from 6e-3 to 9e-2 logits fails.
BF16_MEASURED = {
    "ast_spc2_b8_kr07": ([0.9972, 0.9961, 0.9946], 2.12e-2),
    "ast_spc2_b8_kr07_pert": ([0.9972, 0.9961, 0.9891], 2.11e-2),
    "audiomae_1024_b2_kr07": ([0.9986, 0.9960, 0.9944], 1.96e-2),
    "audiomae_1024_b2_kr07_pert": ([0.9986, 0.9940, 0.9915], 1.44e-2),
    "ast_1024_b2_kr05": ([0.9941, 0.9883, 0.9688], 1.19e-2),
    "ast_1024_b2_kr09_pert": ([0.9989, 0.9976, 0.9987], 5.48e-3),
    "audiomae_256_b3_list": ([1.0, 0.9906, 0.9815, 0.9540], 6.87e-2),
    "audiomae_1024_b4_kr07_trained": ([0.9986, 0.9931, 0.9859], 6.30e-3),
    "ast_1024_b4_kr07_trained": ([0.9965, 0.9921, 0.9915], 5.20e-3),
    "ast_spc2_b4_unpruned": ([], 6.10e-3),
}


def bf16_bounds(name, ks, batch):
    """(min overlap per block, max logits error) from the measured table: two more flipped tokens per block over the
    whole batch than measured, and 1.6x the measured logits error; the north-star 1e-2 wherever it was met."""
    ovs, lg = BF16_MEASURED[name]
    lo = [m - 2.0 / (k * batch) - 1e-4 for m, k in zip(ovs, ks)]
    return lo, (1e-2 if lg < 6.4e-3 else 1.6 * lg)


@pytest.mark.parametrize("precision", ["bf16", "bf16+score32"])
@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_forward_bf16_against_reference_golden(name, precision):
    g = load_golden(name)
    meta, ref, f64 = g["meta"], g["ref"], g["f64"]
    sd, x = make_case(meta)
    model = build_model(meta, sd, precision)
    with torch.no_grad():
        logits, feats = model(x.to(dev()), keep_rate_list=meta["keep_rate_list"], flag_extract_features=True)
    blocks = prune_blocks(ref)
    overlaps = []
    if blocks:
        got, exp = mel_idx(feats, blocks), mel_idx(f64, blocks)
        overlaps = [set_overlap(a, b) for a, b in zip(got, exp)]
    err = rel_err(logits.cpu(), f64["logits"])
    s0 = rel_err(feats["block-0.attn_score"], f64["block-0.attn_score"])
    print(f"[{precision}] {name}: kept-set overlap vs fp64 {['%.4f' % o for o in overlaps]}, logits err {err:.2e}, "
          f"block-0 score err {s0:.2e}")
    # near-uniform scores (random-init weights): 5e-3; peaked attention ("trained" statistics, logits of std ~3): the bf16
    # rounding of q / k moves individual probabilities by a few per cent and the score by < 1.5e-2 of its maximum
    assert s0 < (1.5e-2 if meta["flavour"] == "trained" else 5e-3)
    if blocks:
        # first pruning block sees the same tokens on both sides: mismatches must be near-ties at the cut
        b0 = blocks[0]
        sc = f64[f"block-{b0}.attn_score"]
        kk = f64[f"block-{b0}.topk_idx"].shape[1]
        # ... i.e. closer to the cut than 1 % of it or than twice the largest score error of this block (the flipped
        # token and the token at the cut can each be off by that much)
        max_abs = (feats[f"block-{b0}.attn_score"].double() - sc).abs().max().item()
        for c in range(sc.shape[0]):
            cut = torch.sort(sc[c], descending=True).values[kk - 1].item()
            diff = set(feats[f"block-{b0}.topk_idx"][c].tolist()) ^ set(f64[f"block-{b0}.topk_idx"][c].tolist())
            for tkn in diff:
                assert abs(sc[c, tkn].item() - cut) <= max(1e-2 * abs(cut), 2.0 * max_abs), (b0, c, tkn)
    lo, lg_max = bf16_bounds(name, [e.shape[1] for e in exp] if blocks else [], x.shape[0])
    for o, l in zip(overlaps, lo):
        assert o >= l, (overlaps, lo)
    assert err < lg_max, (err, lg_max)


@pytest.mark.parametrize("name", ["ast_spc2_b4_unpruned", "audiomae_1024_b2_kr07_pert", "audiomae_256_b3_list"])
def test_forward_bf16_with_layernorm_fold(name):
    """The LayerNorm fold (tpat_gemm_ln through tpat_forward, off by default) gives the same forward within the bf16
    tolerances: no separate LayerNorm launch for norm1 of blocks > 0 and norm2 of non-pruning blocks."""
    g = load_golden(name)
    meta, f64 = g["meta"], g["f64"]
    sd, x = make_case(meta)
    plain = build_model(meta, sd, "bf16")
    folded = build_model(meta, sd, "bf16")
    folded._engine.ln_fold = "all"
    with torch.no_grad():
        a = plain(x.to(dev()), keep_rate_list=meta["keep_rate_list"])
        n_plain = plain._engine.last_launch_count
        b = folded(x.to(dev()), keep_rate_list=meta["keep_rate_list"])
        n_fold = folded._engine.last_launch_count
    ia = [t for t in plain.last_topk_idx if t is not None]
    ib = [t for t in folded.last_topk_idx if t is not None]
    same_tokens = all(torch.equal(p.sort(1).values, q.sort(1).values) for p, q in zip(ia, ib))
    ea, eb = rel_err(a.cpu(), f64["logits"]), rel_err(b.cpu(), f64["logits"])
    print(f"[ln fold] {name}: launches {n_plain} -> {n_fold}, logits err vs fp64 {ea:.2e} (plain) {eb:.2e} (folded), same kept sets: {same_tokens}")
    n_prune = len(ia)
    assert n_fold == n_plain - (11 + 12 - n_prune)
    assert eb < 1.6 * max(BF16_MEASURED[name][1], 8.2e-3)


def test_forward_is_deterministic_and_batch_invariant():
    """Full-size property (BASELINE configs[1] shape, B=64): the same clip gives bit-identical
    logits and indices run to run and regardless of which batch it sits in (no atomics, fixed
    reduction orders, per-clip top-k)."""
    from oracle import weights
    meta = dict(variant="audiomae", T=1024, num_classes=527, drop_loc=(3, 6, 9), base_keep_rate=0.7)
    sd = weights.make_audiomae_state_dict(527, 1024, 0, "refinit")
    x = weights.make_spectrogram("audiomae", 64, 1024, 1234).to(dev())
    for precision in ("bf16", "fp32"):
        model = build_model(meta, sd, precision)
        with torch.no_grad():
            a = model(x)
            ia = [t.clone() for t in model.last_topk_idx if t is not None]
            b = model(x)
            ib = [t for t in model.last_topk_idx if t is not None]
            c = model(x[8:16])
            ic = [t for t in model.last_topk_idx if t is not None]
        assert torch.equal(a, b) and all(torch.equal(p, q) for p, q in zip(ia, ib))
        assert torch.equal(a[8:16], c) and all(torch.equal(p[8:16], q) for p, q in zip(ia, ic))
        assert [t.shape[1] for t in ia] == [359, 252, 177]
        assert torch.isfinite(a).all()
        for t, n_in in zip(ia, (512, 359, 252)):   # indices are a duplicate-free subset of the incoming tokens
            assert int(t.min()) >= 0 and int(t.max()) < n_in
            assert all(len(set(r)) == len(r) for r in t[:4].tolist())


def test_cuda_graph_path_matches_eager_launches():
    g = load_golden("ast_spc2_b8_kr07")
    meta = g["meta"]
    sd, x = make_case(meta)
    model = build_model(meta, sd, "bf16")
    with torch.no_grad():
        eager = model(x.to(dev()))
        idx_e = [t.clone() for t in model.last_topk_idx if t is not None]
        model.use_cuda_graph = True
        for _ in range(2):
            graphed = model(x.to(dev()))
        idx_g = [t for t in model.last_topk_idx if t is not None]
    assert torch.equal(eager, graphed) and all(torch.equal(p, q) for p, q in zip(idx_e, idx_g))


def test_model_on_a_second_device_runs_there():
    """One process driving two GPUs (the reference's nn.DataParallel layout): the forward runs on the input's device,
    not on the current one, and gives the same bits as on device 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden("ast_spc2_b8_kr07")
    meta = g["meta"]
    sd, x = make_case(meta)
    m0 = build_model(meta, sd, "bf16")
    m1 = build_model(meta, sd, "bf16").to("cuda:1")
    with torch.no_grad():
        a = m0(x.to("cuda:0"))
        assert torch.cuda.current_device() == 0
        b = m1(x.to("cuda:1"))
    assert b.device.index == 1 and torch.equal(a.cpu(), b.cpu())


# 1.5x the values measured on B200 (profiles/r02*_parity.txt): (logits err, block-3 overlap)
TOL_2048 = {"audiomae": (5e-2, 0.98), "ast": (5e-2, 0.98)}
# bf16, free-running: measured 2.64e-3 (r01g ... r02z) or 3.5e-2 ... 4.0e-2 (r01d, r01f, r02aa: one near-tie flip among the
# 90 / 63 / 45 kept tokens of a clip) for vit_small, 4.33e-2 for vit_large.  The arithmetic itself is held to the
# north-star 1e-2 against the oracle run on the SAME kept tokens (TOL_VIT_FORCED).
TOL_VIT_SIZES = {"vit_small_patch16": 6.5e-2, "vit_large_patch16": 6.5e-2}
TOL_VIT_FORCED = {"vit_small_patch16": 1.0e-2, "vit_large_patch16": 1.0e-2}   # measured 2.50e-3 / 6.68e-3 (r02aa)


@pytest.mark.parametrize("variant", ["audiomae", "ast"])
def test_long_clip_2048_frames(variant):
    """Twice the reference's longest input (2048 frames = 1024 patches, N = 1025 / 1026: nine query tiles, a one- / two-row
    tail tile, 17 key blocks, a 1024-key top-k): fp32 mode against the oracle, and the bf16 mode for finiteness / overlap."""
    from oracle import weights
    T = 2048
    meta = dict(variant=variant, T=T, num_classes=35, drop_loc=(3, 6, 9), base_keep_rate=0.7)
    mk = weights.make_audiomae_state_dict if variant == "audiomae" else weights.make_ast_state_dict
    sd = mk(35, T, 9, "trained")
    x = weights.make_spectrogram(variant, 2, T, 33)
    with torch.no_grad():
        exp, feats = vo.forward(variant, sd, x, None, (3, 6, 9), 0.7)
        m32 = build_model(meta, sd, "fp32")
        got = m32(x.to(dev()))
        i32 = [t.cpu() for t in m32.last_topk_idx if t is not None]
        m16 = build_model(meta, sd, "bf16")
        got16 = m16(x.to(dev()))
        i16 = [t.cpu() for t in m16.last_topk_idx if t is not None]
    assert rel_err(got.cpu(), exp) < 2e-5
    for blk, t in zip((3, 6, 9), i32):
        for a, b in zip(t.tolist(), feats[f"block-{blk}.topk_idx"].tolist()):
            assert set(a) == set(b), blk
    assert [t.shape[1] for t in i16] == [717, 502, 352]
    e16, o16 = rel_err(got16.cpu(), exp), set_overlap(i16[0], i32[0])
    print(f"[bf16 2048 frames] {variant}: logits err {e16:.2e}, block-3 overlap vs fp32 {o16:.4f}")
    assert torch.isfinite(got16).all() and e16 < TOL_2048[variant][0]
    assert o16 > TOL_2048[variant][1]


@pytest.mark.parametrize("variant", ["audiomae", "ast"])
def test_extreme_pruning_down_to_one_token(variant):
    """keep rates that leave 3, then 1, then 1 patch token (128 -> 3 -> 1 -> 1): tiny attention (N = 2 .. 4), top-k with
    k = 1, GEMMs with a handful of rows -- both precisions against the oracle."""
    from oracle import weights
    T = 256
    krl = (1.0, 1.0, 0.02, 1.0, 1.0, 0.3, 1.0, 1.0, 0.5, 1.0, 1.0, 1.0)
    meta = dict(variant=variant, T=T, num_classes=35, drop_loc=(3, 6, 9), base_keep_rate=0.7)
    mk = weights.make_audiomae_state_dict if variant == "audiomae" else weights.make_ast_state_dict
    sd = mk(35, T, 12, "trained")
    x = weights.make_spectrogram(variant, 5, T, 44)
    with torch.no_grad():
        exp, feats = vo.forward(variant, sd, x, krl, (3, 6, 9), 0.7)
        for precision, tol in (("fp32", 2e-5), ("bf16", 5e-2)):
            m = build_model(meta, sd, precision)
            got = m(x.to(dev()), keep_rate_list=krl)
            idx = [t.cpu() for t in m.last_topk_idx if t is not None]
            assert [t.shape[1] for t in idx] == [3, 1, 1]
            if precision == "fp32":
                for blk, t in zip((2, 5, 8), idx):
                    assert torch.equal(t.sort(1).values, feats[f"block-{blk}.topk_idx"].sort(1).values), blk
            assert torch.isfinite(got).all()
            same = all(torch.equal(t.sort(1).values, feats[f"block-{b}.topk_idx"].sort(1).values) for b, t in zip((2, 5, 8), idx))
            if same:
                assert rel_err(got.cpu(), exp) < tol, precision


def test_forward_features_returns_the_classifier_input():
    """models_vit.py:334-396: forward_features = fc_norm(mean of the patch tokens) after the pruned blocks."""
    g = load_golden("audiomae_256_b3_list")
    meta = g["meta"]
    sd, x = make_case(meta)
    model = build_model(meta, sd, "fp32")
    with torch.no_grad():
        feats = model.forward_features(x.to(dev()), keep_rate_list=meta["keep_rate_list"])
        logits = model(x.to(dev()), keep_rate_list=meta["keep_rate_list"])
        _, _, pooled = vo.forward("audiomae", sd, x, meta["keep_rate_list"], meta["drop_loc"], meta["base_keep_rate"],
                                  return_pooled=True)
    assert tuple(feats.shape) == (3, 768)
    assert rel_err(feats.cpu(), pooled) < 2e-5
    assert rel_err((feats @ model.head.weight.T + model.head.bias).cpu(), logits.cpu()) < 1e-5


def test_weights_are_repacked_after_an_update():
    g = load_golden("ast_spc2_b4_unpruned")
    meta = g["meta"]
    sd, x = make_case(meta)
    model = build_model(meta, sd, "bf16")
    with torch.no_grad():
        a = model(x.to(dev()))
        model.mlp_head[1].bias.add_(1.0)               # in-place update bumps the parameter version
        b = model(x.to(dev()))
    assert torch.allclose(b, a + 1.0, atol=1e-5)


@pytest.mark.parametrize("factory,dim,depth,heads", [("vit_small_patch16", 384, 12, 6), ("vit_large_patch16", 1024, 24, 16)])
def test_other_vit_sizes_match_the_oracle(factory, dim, depth, heads):
    """The reference also ships vit_small / vit_large factories (models_vit.py:531-548): same kernels, D = 64 * H."""
    from oracle import weights
    from tpat import models_vit
    T, C, B = 256, 20, 2
    sd = weights.make_audiomae_state_dict(C, T, seed=7, flavour="perturbed", depth=depth, dim=dim)
    x = weights.make_spectrogram("audiomae", B, T, seed=8)
    drop_loc = (1, 3, depth - 2)
    with torch.no_grad():
        ref_logits, ref_feats = vo.forward("audiomae", sd, x, None, drop_loc, 0.7, num_heads=heads)
    for precision, tol in (("fp32", 2e-5), ("bf16", TOL_VIT_SIZES[factory])):
        m = getattr(models_vit, factory)(num_classes=C, drop_path_rate=0.0, mean_pooling=True, mask_2d=True,
                                         target_length=T, drop_loc=drop_loc, base_keep_rate=0.7, precision=precision)
        m.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, dim)
        m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, dim), requires_grad=False)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev()).eval()
        with torch.no_grad():
            logits = m(x.to(dev()))
        err = rel_err(logits.cpu(), ref_logits)
        print(f"[{precision}] {factory}: logits err {err:.2e}")
        assert err < tol
        if precision == "bf16":
            forced = {i: t.cpu() for i, t in enumerate(m.last_topk_idx) if t is not None}
            with torch.no_grad():
                forced_logits, _ = vo.forward("audiomae", sd, x, None, drop_loc, 0.7, num_heads=heads, forced_idx=forced)
            err_f = rel_err(logits.cpu(), forced_logits)
            print(f"[{precision}] {factory}: logits err against the oracle on the same kept tokens {err_f:.2e}")
            assert err_f < TOL_VIT_FORCED[factory]
        if precision == "fp32":
            first = f"block-{drop_loc[0]}.topk_idx"
            for a, b in zip(m.last_topk_idx[drop_loc[0]].cpu().tolist(), ref_feats[first].tolist()):
                assert set(a) == set(b)


@pytest.mark.parametrize("variant,T,kr", [("ast", 1024, 0.5), ("ast", 128, 0.7), ("audiomae", 1024, 0.7), ("ast", 1024, 0.9)])
def test_fused_inattentive_token_matches_the_oracle(variant, T, kr):
    """BASELINE configs[2] names the EViT fused inattentive token; the reference forward does not implement it
    (SURVEY.md F8), so this is checked against the oracle's restatement of upstream EViT -- parity UNPINNED."""
    from oracle import weights
    from tpat import models_vit, ASTModel
    C, B = 35, 2
    if variant == "ast":
        sd = weights.make_ast_state_dict(C, T, seed=11, flavour="perturbed")
        m32 = ASTModel(label_dim=C, input_tdim=T, imagenet_pretrain=False, audioset_pretrain=False, verbose=False,
                       drop_loc=(3, 6, 9), base_keep_rate=kr, precision="fp32", fuse_token=True)
        m32.load_state_dict(sd, strict=False)
    else:
        sd = weights.make_audiomae_state_dict(C, T, seed=11, flavour="perturbed")
        m32 = models_vit.vit_base_patch16(num_classes=C, drop_path_rate=0.0, mean_pooling=True, mask_2d=True, target_length=T,
                                          drop_loc=(3, 6, 9), base_keep_rate=kr, precision="fp32", fuse_token=True)
        m32.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, 768)
        m32.pos_embed = nn.Parameter(torch.zeros(1, m32.patch_embed.num_patches + 1, 768), requires_grad=False)
        m32.load_state_dict(sd, strict=True)
    m32 = m32.to(dev()).eval()
    x = weights.make_spectrogram(variant, B, T, seed=12)
    with torch.no_grad():
        ref_logits, ref = vo.forward(variant, sd, x, None, (3, 6, 9), kr, flag_extract_features=True, fuse_token=True)
        logits, feats = m32(x.to(dev()), flag_extract_features=True)
    assert sorted(k for k in feats if k != "mel") == sorted(ref)
    for k in ref:
        assert feats[k].shape == ref[k].shape, k
    for a, b in zip(feats["block-3.topk_idx"].tolist(), ref["block-3.topk_idx"].tolist()):
        assert set(a) == set(b)
    # block-4 scores include the fused token's.  Exactly tied block-3 scores may be ordered differently than torch.topk
    # does, which permutes the kept tokens, so the kept part is compared as a multiset; the fused
    # token is always last.
    got4, ref4 = feats["block-4.attn_score"], ref["block-4.attn_score"]
    assert rel_err(torch.sort(got4[:, :-1], dim=1).values, torch.sort(ref4[:, :-1], dim=1).values) < 2e-5
    assert rel_err(got4[:, -1], ref4[:, -1]) < 2e-5
    err = rel_err(logits.cpu(), ref_logits)
    print(f"[fp32 fuse] {variant} T={T} kr={kr}: logits err {err:.2e}")
    assert err < 2e-5
    m32.precision = "bf16"
    with torch.no_grad():
        lb = m32(x.to(dev()))
    assert rel_err(lb.cpu(), ref_logits) < 1e-1
