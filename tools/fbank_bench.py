#!/usr/bin/env python
"""Throughput of the GPU log-mel front end (tpat_fbank): 64 clips of 10.24 s at 16 kHz -> [64, 1024, 128], device time
from a CUDA-graph replay of 20 calls; plus the reference pipeline (torchaudio kaldi fbank, oracle/fbank_oracle.py) on the
host cores for a few clips."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "token-pruning-audio-transformer_b200"))
import torch
from tpat.frontend import FbankFrontend
from oracle import fbank_oracle as fo
dev = torch.device("cuda:0")
B, L = 64, 163840
wave = (torch.randn(B, L) * 0.1).to(dev)
fe = FbankFrontend(target_length=1024)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    out = fe(wave); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(20): out = fe(wave)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byt = B * L * 4 + B * 1024 * 128 * 4 * 3      # waveform read + spectrogram written, re-read and re-written by the finalize pass
print(f"tpat_fbank: {ms * 1e3:.1f} us per {B} clips = {B / ms * 1e3:.0f} clips/s, {byt / ms / 1e6:.0f} GB/s algorithmic")
torch.set_num_threads(os.cpu_count() or 1)
x = [fo.make_waveform(L, i) for i in range(4)]
t0 = time.perf_counter()
for w in x: fo.wav2fbank(w)
dt = (time.perf_counter() - t0) / len(x)
print(f"reference pipeline (torchaudio, {torch.get_num_threads()} threads): {dt * 1e3:.1f} ms per clip = {1 / dt:.0f} clips/s")
