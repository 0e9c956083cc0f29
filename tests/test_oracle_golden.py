"""CPU: the oracle restatement against the golden vectors generated from the REAL reference
(oracle/make_golden.py).  Bit-exact: same ATen ops in the same order on the same torch build."""
import pytest
import torch

from conftest import load_golden, make_case
from oracle import vit_oracle as vo
from oracle.golden_configs import ABLATION_CONFIGS, GOLDEN_CONFIGS, MASKED_CONFIGS


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    meta = g["meta"]
    sd, x = make_case(meta)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        logits, feats = vo.forward(meta["variant"], sd, x, meta["keep_rate_list"], meta["drop_loc"],
                                   meta["base_keep_rate"], flag_extract_features=True)
        logits_plain, feats_plain = vo.forward(meta["variant"], sd, x, meta["keep_rate_list"], meta["drop_loc"],
                                               meta["base_keep_rate"], flag_extract_features=False)
    ref = g["ref"]
    # logits: identical up to MKL thread-count effects (bitwise on the generating machine)
    assert torch.allclose(logits, ref["logits"], rtol=0, atol=2e-6)
    assert torch.allclose(logits_plain, g["ref_plain"]["logits"], rtol=0, atol=2e-6)
    keys = sorted(k for k in ref if k != "logits")
    assert sorted(feats) == keys
    for k in keys:
        if k.endswith("topk_idx"):
            # same kept SET per clip; order may differ only between near-tied scores
            for a, b in zip(feats[k].tolist(), ref[k].tolist()):
                assert set(a) == set(b), k
        else:
            assert torch.allclose(feats[k], ref[k], rtol=0, atol=1e-8), k
    # non-extract mode reports only the pruning blocks
    assert sorted(feats_plain) == sorted(k for k in keys if k.split(".")[0] in
                                         {kk.split(".")[0] for kk in keys if kk.endswith("topk_idx")})


def ablation_kwargs(meta):
    return dict(use_custom_rank=meta.get("use_custom_rank"), drop_token_blk_idx=meta.get("drop_token_blk_idx"),
                retain_min=meta.get("retain_min"), retain_max=meta.get("retain_max"))


@pytest.mark.parametrize("name", list(ABLATION_CONFIGS))
def test_oracle_ablation_paths_match_reference_golden(name):
    """custom_rank mean/std and the drop_token_blk_idx intensity filter against the logits the
    real reference produced (models_vit.py:343-385, ast_models.py:445-497)."""
    g = load_golden("abl_" + name)
    meta = g["meta"]
    sd, x = make_case(meta)
    with torch.no_grad():
        logits, info = vo.forward_ablation(meta["variant"], sd, x, meta["keep_rate_list"], meta["drop_loc"],
                                           meta["base_keep_rate"], **ablation_kwargs(meta))
    ref = g["ref"]["logits"]
    if ref is None:
        assert logits is None                       # nothing retained -> the reference returns None
        return
    assert torch.allclose(logits, ref, rtol=0, atol=2e-6)
    for blk, idx in g["oracle_info"]["topk_idx"].items():
        for a, b in zip(info["topk_idx"][blk].tolist(), idx.tolist()):
            assert set(a) == set(b), blk
    if g["oracle_info"]["retain_idx"] is not None:
        assert torch.equal(info["retain_idx"], g["oracle_info"]["retain_idx"])


@pytest.mark.parametrize("name", list(MASKED_CONFIGS))
def test_oracle_masked_forward_matches_reference_golden(name):
    """Forward half of the fine-tune 2-D masking (models_vit.py:425-497) against the reference."""
    g = load_golden("mask_" + name)
    meta = g["meta"]
    sd, x = make_case(meta)
    keep_idx = vo.masking_2d_keep_indices(g["noise_t"], g["noise_f"], meta["mask_t_prob"], meta["mask_f_prob"])
    assert torch.equal(keep_idx, g["keep_idx"])
    with torch.no_grad():
        logits = vo.forward_masked(meta["variant"], sd, x, keep_idx, meta["keep_rate_list"], meta["drop_loc"],
                                   meta["base_keep_rate"])
    assert torch.allclose(logits, g["ref"]["logits"], rtol=0, atol=2e-6)


def test_oracle_custom_rank_keeps_exactly_k_tokens():
    """The reference's custom-rank gather indexes the full token list, so a pruning block hands on k tokens (no
    separately kept cls row): 1 + 128 -> 90 -> 63 -> 44 at 256 frames, keep 0.7."""
    g = load_golden("abl_audiomae_256_b2_rank_mean")
    assert [g["oracle_info"]["topk_idx"][b].shape[1] for b in (3, 6, 9)] == [90, 63, 44]


def test_fp64_oracle_agrees_with_fp32_reference_sets():
    g = load_golden("audiomae_1024_b2_kr07")
    for k, v in g["ref"].items():
        if k.endswith("topk_idx"):
            for a, b in zip(v.tolist(), g["f64"][k].tolist()):
                assert len(set(a) & set(b)) >= len(a) - 1, k   # fp32 vs fp64: at most one near-tie flip


def test_token_schedule_matches_survey():
    # 512 -> 359 -> 252 -> 177 ; SPC-2: 64 -> 45 -> 32 -> 23
    rates = vo.default_keep_rate_list(12, (3, 6, 9), 0.7)
    assert vo.token_schedule(512, rates)[3::3] == [359, 252, 177]
    assert vo.token_schedule(64, rates)[3::3] == [45, 32, 23]
    assert vo.token_schedule(512, vo.default_keep_rate_list(12, (3, 6, 9), 0.5))[3::3] == [256, 128, 64]


def test_melspec_indices_composition():
    a = torch.tensor([[3, 1, 2]])
    b = torch.tensor([[2, 0]])
    out = vo.melspec_indices([a, b])
    assert out[1].tolist() == [[2, 3]]


def test_fbank_oracle_matches_golden_and_mel_table_matches_torchaudio():
    """The front-end oracle (torchaudio kaldi fbank + the loaders' pad / crop / normalise) against the committed
    outputs, and the host-built mel filter table against torchaudio's (bit-identical)."""
    pytest.importorskip("torchaudio")
    import torchaudio.compliance.kaldi as kaldi
    from oracle import fbank_oracle as fo
    from tpat.frontend import FbankFrontend, mel_banks
    g = load_golden("fbank_cases")
    for name in g["meta"]["cases"]:
        c = g[name]
        spec = fo.wav2fbank(fo.make_waveform(c["n"], c["seed"]), target_length=c["T"])
        assert torch.allclose(spec, c["spec"], rtol=0, atol=2e-5), name
    ref, _ = kaldi.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    mine, start, length = mel_banks(128, 512, 16000.0)
    assert torch.equal(mine[:, :256], ref) and torch.all(mine[:, 256] == 0)
    for m in range(128):
        nz = torch.nonzero(mine[m] > 0).flatten()
        if nz.numel():
            assert int(start[m]) == int(nz[0]) and int(start[m] + length[m] - 1) == int(nz[-1])
    fe = FbankFrontend()
    assert (fe.win, fe.shift, fe.nfft, fe.num_frames(163840)) == (400, 160, 512, 1022)


# ---- fine-tune step: the oracle's autograd against the REAL reference's gradients (tests/golden/grad_*.pt) ----------

def _grad_case(cfg):
    import hashlib
    from oracle import weights
    mk = weights.make_audiomae_state_dict if cfg["variant"] == "audiomae" else weights.make_ast_state_dict
    sd = mk(cfg["num_classes"], cfg["T"], cfg["wseed"], cfg["flavour"])
    x = weights.make_spectrogram(cfg["variant"], cfg["B"], cfg["T"], cfg["xseed"])
    assert weights.state_dict_digest(sd) == cfg["sd_digest"] and hashlib.sha256(x.numpy().tobytes()).hexdigest() == cfg["x_digest"]
    y = (torch.rand(cfg["B"], cfg["num_classes"], generator=torch.Generator().manual_seed(cfg["tseed"])) < 0.1).float()
    return sd, x, y


@pytest.mark.parametrize("name", ["audiomae_256_b2_train", "audiomae_256_b2_train_masked", "ast_128_b2_train"])
def test_oracle_autograd_reproduces_reference_gradients(name):
    """Train-mode forward + backward of the restatement (DropPath draws and 2-D masking noise replayed in the reference's
    RNG order) against the gradient fingerprints of the real reference: loss, logits, and for every parameter the l2
    norm and 64 strided samples."""
    g = load_golden("grad_" + name)
    cfg = g["meta"]
    sd, x, y = _grad_case(cfg)
    torch.manual_seed(cfg["dseed"])
    keep_idx = None
    if cfg["mask_t_prob"] > 0 or cfg["mask_f_prob"] > 0:
        noise_t, noise_f = torch.rand(cfg["B"], cfg["T"] // 16), torch.rand(cfg["B"], 8)
        keep_idx = vo.masking_2d_keep_indices(noise_t, noise_f, cfg["mask_t_prob"], cfg["mask_f_prob"])
        assert torch.equal(keep_idx, g["keep_idx"])
    scales = vo.drop_path_scales(cfg["drop_rates"], cfg["B"])
    frozen = ("pos_embed",) if cfg["variant"] == "audiomae" else ()
    loss, logits, grads = vo.loss_and_grads(cfg["variant"], sd, x, y, cfg["keep_rate_list"], cfg["drop_loc"], cfg["base_keep_rate"],
                                            dtype=torch.float32, drop_scales=scales, mask_keep_idx=keep_idx, frozen=frozen)
    assert abs(loss.item() - g["loss"]) < 1e-6
    assert torch.allclose(logits, g["logits"], rtol=0, atol=2e-6 * g["logits"].abs().max().item())
    assert sorted(grads) == sorted(g["grads"])
    for k, summ in g["grads"].items():
        f = grads[k].reshape(-1).double()
        stride = max(1, f.numel() // 64)
        assert f.numel() == summ["numel"]
        assert abs(f.norm().item() - summ["norm"]) <= 1e-5 * summ["norm"] + 1e-12, k
        assert ((f[::stride][:64] - summ["samples"]).norm() / summ["samples"].norm().clamp_min(1e-30)).item() < 1e-4, k


def test_oracle_gradient_structure():
    """No gradient through the score / top-k; dropped tokens still feed the patch-embed gradient through earlier blocks;
    a clip whose DropPath mask is 0 in every block contributes nothing to the block weights."""
    from oracle import weights
    sd = weights.make_audiomae_state_dict(5, 128, 3, "perturbed", depth=2)
    x = weights.make_spectrogram("audiomae", 2, 128, 4)
    y = torch.zeros(2, 5); y[:, 1] = 1
    drops = [(torch.tensor([0.0, 2.0]), torch.tensor([0.0, 2.0]))] * 2
    _, _, g = vo.loss_and_grads("audiomae", sd, x, y, (0.5, 1.0), (0,), 0.5, num_heads=12, dtype=torch.float64, drop_scales=drops)
    assert "pos_embed" not in g and len(g) == len([k for k in sd if k != "pos_embed"])
    assert all(torch.isfinite(v).all() for v in g.values())
