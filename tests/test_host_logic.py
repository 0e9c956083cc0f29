"""CPU: the host-side mirror of the reference model API (names, state-dict keys, errors, schedule)."""
import math
import os

import pytest
import torch
import torch.nn as nn

import conftest  # noqa: F401
from oracle import vit_oracle as vo, weights


def build_audiomae(T=1024, C=527, **kw):
    from tpat import models_vit
    m = models_vit.vit_base_patch16(num_classes=C, drop_path_rate=0.1, mean_pooling=True, mask_2d=True,
                                    target_length=T, drop_loc=(3, 6, 9), base_keep_rate=0.7, **kw)
    m.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, 768)             # main_finetune.py:378
    m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, 768), requires_grad=False)
    return m


def test_audiomae_state_dict_keys_match_reference_format():
    m = build_audiomae()
    sd = weights.make_audiomae_state_dict(527, 1024, 0)
    assert set(m.state_dict()) == set(sd)
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == sum(v.numel() for v in sd.values())
    assert len(m.blocks) == 12 and m.no_weight_decay() == {'pos_embed', 'cls_token'}
    assert [b.attn.default_keep_rate for b in m.blocks] == vo.default_keep_rate_list(12, (3, 6, 9), 0.7)


def test_ast_state_dict_keys_match_reference_format():
    from tpat import ASTModel
    m = ASTModel(label_dim=35, input_tdim=128, imagenet_pretrain=False, audioset_pretrain=False, verbose=False,
                 drop_loc=(3, 6, 9), base_keep_rate=0.7)
    sd = weights.make_ast_state_dict(35, 128, 0)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    assert set(missing) == {"v.head.weight", "v.head.bias", "v.head_dist.weight", "v.head_dist.bias"}
    assert m.v.pos_embed.shape == (1, 66, 768)
    assert [b.attn.default_keep_rate for b in m.v.blocks] == vo.default_keep_rate_list(12, (3, 6, 9), 0.7)
    assert all(b.attn.num_extra_tokens == 2 and b.num_extra_tokens == 2 for b in m.v.blocks)
    # DataParallel-style checkpoints carry a "module." prefix (traintest.py:247); the mirror's keys are the suffixes
    assert all(k.startswith(("v.", "mlp_head.")) for k in m.state_dict())


def test_errors_match_reference_behaviour():
    m = build_audiomae(T=128, C=10)
    with pytest.raises(ValueError):                      # models_vit.py:506-507
        m(torch.zeros(1, 1, 128, 128), keep_rate_list=[1.0] * 11)
    with pytest.raises(RuntimeError):                    # no CPU path
        m(torch.zeros(1, 1, 128, 128))
    with pytest.raises(AssertionError):                  # models_vit.py:336: T >= F and F == 128
        m(torch.zeros(1, 1, 64, 128))
    with pytest.raises(RuntimeError):                    # masked / ablation paths have no CPU path either
        m(torch.zeros(1, 1, 128, 128), mask_t_prob=0.3)
    m.use_custom_rank = "median"
    with pytest.raises((ValueError, RuntimeError)):      # unknown statistic (ast_models.py:453) / no CPU path
        m(torch.zeros(1, 1, 128, 128))
    m.use_custom_rank = None
    from tpat import models_vit
    with pytest.raises(AssertionError):                  # models_vit.py:66
        models_vit.Attention(768, 12, default_keep_rate=0.0)


def test_masking_indices_follow_reference_construction():
    """random_masking_2d (models_vit.py:425-463) restated as an index list: compare with the reference's gathers on
    a token-id grid, same noise."""
    m = build_audiomae(T=256, C=10)
    gen = torch.Generator().manual_seed(11)
    noise = (torch.rand(2, 16, generator=gen), torch.rand(2, 8, generator=gen))
    keep_idx = m.random_masking_2d_indices(2, torch.device("cpu"), 0.3, 0.25, noise=noise)
    assert tuple(keep_idx.shape) == (2, int(16 * 0.7) * int(8 * 0.75))
    tok = torch.arange(128).reshape(1, 16, 8).expand(2, -1, -1)
    ids_t = torch.argsort(noise[0], dim=1)[:, :11]
    ids_f = torch.argsort(noise[1], dim=1)[:, :6]
    t1 = torch.gather(tok, 1, ids_t[:, :, None].expand(-1, -1, 8)).permute(0, 2, 1)           # N F T'
    t2 = torch.gather(t1, 1, ids_f[:, :, None].expand(-1, -1, 11)).permute(0, 2, 1).reshape(2, -1)
    assert torch.equal(keep_idx, t2)
    assert torch.equal(keep_idx, vo.masking_2d_keep_indices(noise[0], noise[1], 0.3, 0.25))


def test_keep_rate_schedule_matches_reference_golden():
    """engine_finetune.py:29-53: all-ones before the shrink phase, half-cosine descent during it, None afterwards;
    golden values produced by the reference function itself (tests/golden/keep_rate_schedule.pt)."""
    from conftest import load_golden
    from tpat.schedule import get_scheduled_keep_rate_list
    for kwargs, want in load_golden("keep_rate_schedule")["cases"]:
        got = get_scheduled_keep_rate_list(**kwargs)
        assert got == want, (kwargs, got, want)


def test_precision_selection():
    assert build_audiomae(T=128, C=10).precision == "bf16"
    assert build_audiomae(T=128, C=10, precision="fp32").precision == "fp32"
    with pytest.raises(ValueError):
        build_audiomae(T=128, C=10, precision="fp16")


@pytest.mark.parametrize("n,extra", [(512, 1), (512, 2), (64, 2), (128, 1)])
@pytest.mark.parametrize("kr", [0.5, 0.6, 0.7, 0.8, 0.9, 0.999, 1.0])
def test_pruning_schedule_follows_reference_ceil(n, extra, kr):
    from tpat.engine import pruning_schedule
    rates = vo.default_keep_rate_list(12, (3, 6, 9), kr)
    prune, keep = pruning_schedule(n, extra, rates)
    assert keep == vo.token_schedule(n, rates)
    assert prune == [1 if r < 1.0 else 0 for r in rates]          # top-k runs whenever keep_rate < 1.0
    cur = n
    for i in (3, 6, 9):
        if kr < 1.0:
            assert keep[i] == math.ceil(kr * cur)
            cur = keep[i]


def test_shard_bounds_cover_batch():
    from tpat.dist import shard_bounds
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pruning_schedule_with_fused_token():
    """EViT fused token (not in the reference forward): a block that drops tokens hands k + 1 tokens on."""
    from tpat.engine import pruning_schedule, tokens_entering
    rates = vo.default_keep_rate_list(12, (3, 6, 9), 0.7)
    prune, keep = pruning_schedule(512, 2, rates, fuse_token=True)
    assert keep[3] == 359 and keep[6] == math.ceil(0.7 * 360) and keep[9] == math.ceil(0.7 * (keep[6] + 1))
    ent = tokens_entering(512, prune, keep, True)
    assert ent[:4] == [512] * 4 and ent[4] == 360 and ent[7] == keep[6] + 1 and ent[10] == keep[9] + 1
    # keep_rate < 1 that drops nothing (k == n) appends no fused token
    prune, keep = pruning_schedule(4, 1, [0.99] + [1.0] * 11, fuse_token=True)
    assert prune[0] == 1 and keep[0] == 4 and tokens_entering(4, prune, keep, True)[1] == 4


def test_oracle_fused_token_shapes():
    sd = weights.make_ast_state_dict(35, 128, 0, "perturbed")
    x = weights.make_spectrogram("ast", 2, 128, 3)
    with torch.no_grad():
        logits, feats = vo.forward("ast", sd, x, None, (3, 6, 9), 0.7, flag_extract_features=True, fuse_token=True)
    assert logits.shape == (2, 35)
    assert feats["block-3.topk_idx"].shape == (2, 45) and feats["block-4.attn_score"].shape == (2, 46)
    assert feats["block-6.topk_idx"].shape == (2, math.ceil(0.7 * 46))


def test_extract_contract_helpers(tmp_path):
    """N2: the on-disk / index-composition contract extract_stats.py consumes."""
    from tpat import extract
    g = torch.Generator().manual_seed(0)
    i3 = torch.stack([torch.randperm(64, generator=g)[:45] for _ in range(2)])
    i6 = torch.stack([torch.randperm(45, generator=g)[:32] for _ in range(2)])
    i9 = torch.stack([torch.randperm(32, generator=g)[:23] for _ in range(2)])
    feats = {"mel": torch.randn(2, 1, 128, 128, generator=g), "block-3.attn_score": torch.rand(2, 64),
             "block-3.topk_idx": i3, "block-6.topk_idx": i6, "block-9.topk_idx": i9}
    paths = extract.save_feature_dict(feats, str(tmp_path), 7)
    assert sorted(os.path.basename(p) for p in paths) == sorted(f"{k}.0007.pth" for k in feats)
    assert torch.equal(torch.load(str(tmp_path / "block-6.topk_idx.0007.pth")), i6)
    mel = extract.get_melspec_idx(extract.topk_indices_of(feats))
    assert [t.tolist() for t in mel] == [t.tolist() for t in vo.melspec_indices([i3, i6, i9])]
    assert torch.equal(mel[2][0], i3[0][i6[0][i9[0]]])
    # fused-token variant: index len(prev) maps to the dummy 0 column
    f6 = i6.clone(); f6[:, 0] = 45
    assert extract.get_melspec_idx([i3, f6], fuse_token=True)[1][:, 0].tolist() == [0, 0]
    # apply_mask keeps exactly the listed 16x16 patches
    x = feats["mel"]
    masked = extract.apply_mask(x, mel[2])
    patches = x.reshape(2, 1, 8, 16, 8, 16)
    keep = torch.zeros(2, 64, dtype=torch.bool); keep.scatter_(1, mel[2], True)
    ref = (patches * keep.reshape(2, 1, 8, 1, 8, 1)).reshape(2, 1, 128, 128)
    assert torch.equal(masked, ref)


def test_llrd_groups_match_reference_rule():
    """tpat.lr_decay.param_groups_lrd: layer ids, scales and decay flags as util/lr_decay.py:15-75 assigns them."""
    from tpat.lr_decay import get_layer_id_for_vit, param_groups_lrd
    m = build_audiomae(T=128, C=10)
    m.pos_embed.requires_grad_(False)
    groups = param_groups_lrd(m, 0.05, no_weight_decay_list=m.no_weight_decay(), layer_decay=0.75)
    assert len(groups) == 28
    names = {id(p): n for n, p in m.named_parameters()}
    for g in groups:
        ids = {get_layer_id_for_vit(names[id(p)], 13) for p in g["params"]}
        assert len(ids) == 1
        lid = ids.pop()
        assert abs(g["lr_scale"] - 0.75 ** (13 - lid)) < 1e-12
        for p in g["params"]:
            no_decay = p.ndim == 1 or names[id(p)] in ("cls_token", "pos_embed")
            assert g["weight_decay"] == (0.0 if no_decay else 0.05)
    assert get_layer_id_for_vit("cls_token", 13) == 0 and get_layer_id_for_vit("blocks.11.mlp.fc2.bias", 13) == 12
    assert get_layer_id_for_vit("fc_norm.weight", 13) == 13 and get_layer_id_for_vit("head.bias", 13) == 13
    assert sum(len(g["params"]) for g in groups) == 151


def test_train_engine_flattens_parameters_by_backward_stage():
    """TrainEngine.attach (no GPU needed): p.data become views of one flat buffer ordered head -> blocks -> patch embed,
    each stage's gradients are one contiguous slice (the all-reduce buckets), load_state_dict keeps the views."""
    from tpat.train import TrainEngine
    from tpat import _lib
    m = build_audiomae(T=128, C=10)
    m.pos_embed.requires_grad_(False)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    eng = TrainEngine(_lib.VARIANT_AUDIOMAE, 12, 768, 12, 3072)
    roles = m._train_roles()
    eng.attach(m._train_entries(roles), {})
    assert len(eng.entries) == 151 and eng.flat_p.numel() % 64 == 0
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    base = eng.flat_p.data_ptr()
    for _, _, p in eng.entries:
        off, n = eng.offsets[id(p)]
        assert p.data_ptr() == base + 4 * off and off % 64 == 0
    spans = [eng.stage_slices[s] for s in range(13, -1, -1)]
    assert spans[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][0] + spans[-1][1] == eng.flat_p.numel()
    flat_id = eng.flat_p.data_ptr()
    eng.attach(m._train_entries(m._train_roles()), {})           # second call: nothing to rebuild
    assert eng.flat_p.data_ptr() == flat_id
    m.load_state_dict(before)                                     # copies INTO the views
    assert m.head.weight.data_ptr() == base + 4 * eng.offsets[id(m.head.weight)][0]
    g = eng.grad_view(m.head.bias)
    assert g.shape == m.head.bias.shape and g.data_ptr() == eng.flat_g.data_ptr() + 4 * eng.offsets[id(m.head.bias)][0]
