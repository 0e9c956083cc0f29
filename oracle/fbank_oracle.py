"""TEST INFRASTRUCTURE ONLY.  The reference's eval-time input pipeline restated around its pinned third-party
dependency: torchaudio.compliance.kaldi.fbank (torchaudio==2.4.1 in amae_pruning_miniconda.yml:90; 2.11 here -- the
kaldi-compliance module is unchanged between the two).  Follows audiomae/dataset.py:175-178,209-229,298 and
ast/src/dataloader.py:98-101,129-147,204.  Never imported by the product path.
"""
import torch


def wav2fbank(waveform: torch.Tensor, sr: int = 16000, melbins: int = 128, target_length: int = 1024,
              norm_mean: float = -4.2677393, norm_std: float = 4.5689974) -> torch.Tensor:
    """waveform [1, n] fp32 -> [target_length, melbins] fp32 (single-clip path, no mixup / SpecAug)."""
    import torchaudio
    waveform = waveform - waveform.mean()                                         # dataset.py:178
    fbank = torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=sr, use_energy=False,
                                              window_type='hanning', num_mel_bins=melbins, dither=0.0,
                                              frame_shift=10)                      # :209-210
    n_frames = fbank.shape[0]
    p = target_length - n_frames
    if p > 0:                                                                      # :217-221
        fbank = torch.nn.ConstantPad2d((0, 0, 0, p), fbank.min())(fbank)
    elif p < 0:
        fbank = fbank[0:target_length, :]                                          # :223
    return (fbank - norm_mean) / (norm_std * 2)                                    # :298


def make_waveform(n: int, seed: int, sr: int = 16000) -> torch.Tensor:
    """Deterministic test signal: two chirping tones + noise with a slow envelope, amplitude ~0.3, [1, n]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64) / sr
    f1, f2 = 200.0 + 50.0 * seed, 1800.0 + 300.0 * seed
    x = 0.2 * torch.sin(2 * torch.pi * (f1 * t + 40.0 * t * t)) + 0.1 * torch.sin(2 * torch.pi * f2 * t)
    env = 0.6 + 0.4 * torch.sin(2 * torch.pi * 0.7 * t)
    x = x * env + 0.03 * torch.randn(n, generator=g, dtype=torch.float64) + 0.01
    return x.to(torch.float32).unsqueeze(0)
