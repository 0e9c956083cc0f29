"""GPU front end: waveform batch -> normalised log-mel spectrogram (the model's input), computed by libtpat.so.

Mirrors the eval path of the reference data loaders (audiomae/dataset.py:175-230,298; ast/src/dataloader.py:98-149,204):

    waveform = waveform - waveform.mean()
    fbank = torchaudio.compliance.kaldi.fbank(waveform, htk_compat=True, sample_frequency=sr, use_energy=False,
                                              window_type='hanning', num_mel_bins=128, dither=0.0, frame_shift=10)
    pad with fbank.min() / crop to target_length frames;  fbank = (fbank - norm_mean) / (norm_std * 2)

Only the small constant tables (Hann window, triangular mel filters) are built here, with the same fp32 formulae
torchaudio uses (kaldi.py: _feature_window_function, get_mel_banks); every per-sample operation runs in the
``tpat_fbank`` kernels.  Augmentations (mixup, SpecAug, roll) are training-time host code and stay out of scope.
"""
import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, lib

# dataset statistics the reference passes as norm_mean / norm_std (main_finetune.py:252-254; ast run.py)
NORM_STATS = {"audioset": (-4.2677393, 4.5689974), "esc50": (-6.6268077, 5.358466), "spc2": (-6.845978, 5.5654526)}


def mel_banks(num_bins: int, nfft: int, sample_freq: float, low_freq: float = 20.0, high_freq: float = 0.0):
    """kaldi.get_mel_banks (vtln_warp 1.0) + the zero column fbank() appends: ([num_bins, nfft/2+1] fp32, start, len)."""
    num_fft_bins = nfft // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / nfft
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up, down = (mel - left) / (center - left), (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    bins = torch.nn.functional.pad(bins, (0, 1), mode="constant", value=0).contiguous()
    nz = bins > 0
    start = torch.where(nz.any(1), nz.float().argmax(1), torch.zeros(num_bins, dtype=torch.long))
    last = bins.shape[1] - 1 - nz.flip(1).float().argmax(1)
    length = torch.where(nz.any(1), last - start + 1, torch.zeros(num_bins, dtype=torch.long))
    return bins, start.to(torch.int32), length.to(torch.int32)


class FbankFrontend:
    """``frontend(wave [B, L] fp32 CUDA, lengths=None) -> spec [B, target_length, num_mel_bins] fp32``."""

    def __init__(self, sample_rate: int = 16000, num_mel_bins: int = 128, target_length: int = 1024,
                 norm_mean: float = NORM_STATS["audioset"][0], norm_std: float = NORM_STATS["audioset"][1],
                 frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0, preemphasis: float = 0.97,
                 subtract_clip_mean: bool = True):
        self.sr = sample_rate
        self.n_mel, self.T = num_mel_bins, target_length
        self.norm_mean, self.norm_std = float(norm_mean), float(norm_std)
        self.win = int(sample_rate * frame_length_ms * 0.001)                 # kaldi.py:_get_waveform_and_window_properties
        self.shift = int(sample_rate * frame_shift_ms * 0.001)
        self.nfft = 1 << (self.win - 1).bit_length()                          # round_to_power_of_two
        self.preemph = float(preemphasis)
        self.subtract_clip_mean = bool(subtract_clip_mean)
        self._window = torch.hann_window(self.win, periodic=False, dtype=torch.float32)
        self._mel, self._mstart, self._mlen = mel_banks(num_mel_bins, self.nfft, float(sample_rate))
        self._dev = None

    def _tables(self, device):
        if self._dev != device:
            self._w_d, self._mel_d = self._window.to(device), self._mel.to(device)
            self._ms_d, self._ml_d = self._mstart.to(device), self._mlen.to(device)
            self._dev = device
        return self._w_d, self._mel_d, self._ms_d, self._ml_d

    def num_frames(self, n_samples: int) -> int:
        return 1 + (n_samples - self.win) // self.shift if n_samples >= self.win else 0

    def __call__(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None
                 ) -> torch.Tensor:
        if not wave.is_cuda:
            raise RuntimeError("tpat: the front end takes CUDA tensors; there is no CPU path")
        if wave.dim() != 2 or wave.dtype != torch.float32 or not wave.is_contiguous():
            raise RuntimeError("wave must be a contiguous fp32 [B, L] tensor (mono)")
        with torch.cuda.device(wave.device):
            return self._call(wave, lengths, out)

    def _call(self, wave: torch.Tensor, lengths: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, L = wave.shape
        w, mel, ms, ml = self._tables(wave.device)
        if lengths is not None:
            lengths = lengths.to(device=wave.device, dtype=torch.int32).contiguous()
        if out is not None:      # caller-owned spectrogram buffer (a fixed address keeps a static-IO CUDA graph valid)
            if tuple(out.shape) != (B, self.T, self.n_mel) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != wave.device:
                raise RuntimeError("out must be a contiguous fp32 [B, target_length, num_mel_bins] tensor on the input's device")
            spec = out
        else:
            spec = torch.empty(B, self.T, self.n_mel, device=wave.device, dtype=torch.float32)
        ws = torch.empty(32 * B, device=wave.device, dtype=torch.float32)
        check(lib.tpat_fbank(wave.data_ptr(), None if lengths is None else lengths.data_ptr(), B, L,
                             1 if self.subtract_clip_mean else 0, ws.data_ptr(), w.data_ptr(), self.win,
                             self.shift, self.nfft, self.preemph, mel.data_ptr(), ms.data_ptr(), ml.data_ptr(), self.n_mel,
                             spec.data_ptr(), self.T, self.norm_mean, self.norm_std, torch.cuda.current_stream().cuda_stream),
              "tpat_fbank")
        return spec
