"""GPU parity of the log-mel front end against the reference's own pipeline: torchaudio's kaldi fbank
(the pinned third-party dependency the data loaders call) + pad-with-minimum / crop + normalisation, restated in
oracle/fbank_oracle.py.  Floating point: fp32 FFT of different summation order -> tolerance 2e-4 absolute on the
normalised log-mel values (range about -1.3 .. 1), 1e-5 on average; written here, measured values printed."""
import pytest
import torch

import conftest  # noqa: F401
from conftest import load_golden
from gpu_util import dev
from oracle import fbank_oracle as fo

pytestmark = pytest.mark.gpu


def run_frontend(waves, T, lengths=None, **kw):
    from tpat.frontend import FbankFrontend
    fe = FbankFrontend(target_length=T, **kw)
    L = max(w.shape[1] for w in waves)
    batch = torch.zeros(len(waves), L)
    for i, w in enumerate(waves):
        batch[i, :w.shape[1]] = w[0]
    lens = None if lengths is None else torch.tensor(lengths, dtype=torch.int32)
    return fe(batch.to(dev()), lens).cpu()


@pytest.mark.parametrize("n,T", [(16000, 128), (32960, 204), (24000, 100), (163840, 1024), (400, 16)])
def test_fbank_matches_torchaudio_pipeline(n, T):
    x = fo.make_waveform(n, seed=n % 7)
    got = run_frontend([x], T)[0]
    exp = fo.wav2fbank(x, target_length=T)
    err = (got - exp).abs()
    print(f"[fbank] n={n} T={T}: max abs err {err.max():.2e}, mean {err.mean():.2e}")
    assert tuple(got.shape) == (T, 128)
    assert err.max() < 2e-4 and err.mean() < 1e-5


def test_fbank_ragged_batch_pads_each_clip_with_its_own_minimum():
    lens = [163840, 48000, 16000, 100000]
    waves = [fo.make_waveform(n, seed=i + 1) for i, n in enumerate(lens)]
    got = run_frontend(waves, 1024, lengths=lens)
    for i, w in enumerate(waves):
        exp = fo.wav2fbank(w, target_length=1024)
        assert (got[i] - exp).abs().max() < 2e-4, i
        nf = 1 + (lens[i] - 400) // 160
        if nf < 1024:                                   # padded rows are constant = the clip's normalised minimum
            assert torch.all(got[i, nf:] == got[i, nf:].flatten()[0])
            assert abs(got[i, nf, 0].item() - got[i, :nf].min().item()) < 1e-6


def test_fbank_golden_cases():
    g = load_golden("fbank_cases")
    for name in g["meta"]["cases"]:
        c = g[name]
        x = fo.make_waveform(c["n"], c["seed"])
        got = run_frontend([x], c["T"])[0]
        assert (got - c["spec"]).abs().max() < 2e-4, name


def test_frontend_feeds_the_model_like_the_reference_loader():
    """waveform -> tpat_fbank -> tpat_forward equals oracle fbank -> oracle forward (fp32 mode), two clips."""
    from oracle import vit_oracle as vo, weights
    from test_gpu_20_forward import build_model
    from gpu_util import rel_err
    meta = dict(variant="audiomae", T=256, num_classes=35, drop_loc=(3, 6, 9), base_keep_rate=0.7)
    sd = weights.make_audiomae_state_dict(35, 256, 3, "perturbed")
    waves = [fo.make_waveform(256 * 160 + 240, seed=s) for s in (4, 5)]
    spec = run_frontend(waves, 256)
    exp_spec = torch.stack([fo.wav2fbank(w, target_length=256) for w in waves])
    model = build_model(meta, sd, "fp32")
    with torch.no_grad():
        got = model(spec.unsqueeze(1).to(dev()))
        exp, _ = vo.forward("audiomae", sd, exp_spec.unsqueeze(1), None, (3, 6, 9), 0.7)
    assert rel_err(got.cpu(), exp) < 1e-4
