"""Log-mel front-end on the GPU — drop-in for the reference's ``to_mel`` / ``preprocess``
(``test.py:40-47``, ``meldataset.py:42-49``): ``torchaudio.transforms.MelSpectrogram(n_mels=80,
n_fft=2048, win_length=1200, hop_length=300)`` followed by ``(log(1e-5 + mel) + 4) / 4``.

The transform's constants (periodic Hann window zero-padded to ``n_fft``, HTK triangular filterbank with
torchaudio's default ``sample_rate=16000`` — the reference never passes its 24 kHz rate, which is kept for
checkpoint compatibility — and the FFT twiddles) are computed once on the host; the arithmetic is one
launch of ``as_log_mel`` (``csrc/frontend.cu``) for a whole batch of recordings.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops

N_FFT, WIN_LENGTH, HOP_LENGTH, N_MELS = 2048, 1200, 300, 80
LOG_EPS, MEAN, STD = 1e-5, -4.0, 4.0


def hann_window_padded(n_fft: int = N_FFT, win_length: int = WIN_LENGTH) -> torch.Tensor:
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    out = torch.zeros(n_fft, dtype=torch.float64)
    left = (n_fft - win_length) // 2
    out[left:left + win_length] = w
    return out.float()


def htk_filterbank(n_freqs: int, n_mels: int, sample_rate: int = 16000) -> torch.Tensor:
    """Triangular HTK mel filters without normalisation, ``[n_freqs, n_mels]`` (f_min 0, f_max sample_rate/2)."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    m_pts = torch.linspace(to_mel(0.0), to_mel(sample_rate / 2.0), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def filter_ranges(fb: torch.Tensor) -> torch.Tensor:
    """int32 ``[n_mels, 2]``: first bin and number of bins with a non-zero weight per filter."""
    out = torch.zeros(fb.shape[1], 2, dtype=torch.int32)
    for m in range(fb.shape[1]):
        nz = torch.nonzero(fb[:, m] > 0).flatten()
        if nz.numel():
            out[m, 0] = int(nz[0])
            out[m, 1] = int(nz[-1]) - int(nz[0]) + 1
    return out


class LogMel(nn.Module):
    """``forward(wave [B, N] fp32 CUDA, lengths=None) -> [B, 80, 1 + N // 300]`` normalised log-mel."""

    def __init__(self, n_mels: int = N_MELS, n_fft: int = N_FFT, win_length: int = WIN_LENGTH,
                 hop_length: int = HOP_LENGTH, mean: float = MEAN, std: float = STD):
        super().__init__()
        if n_fft != 2048:
            raise NotImplementedError("as_log_mel implements the reference's n_fft = 2048")
        self.n_mels, self.n_fft, self.hop_length, self.mean, self.std = n_mels, n_fft, hop_length, mean, std
        fb = htk_filterbank(n_fft // 2 + 1, n_mels)
        k = np.arange(n_fft // 2, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / n_fft), -np.sin(2 * np.pi * k / n_fft)], axis=1)
        self.register_buffer("window", hann_window_padded(n_fft, win_length), persistent=False)
        self.register_buffer("fb", fb.contiguous(), persistent=False)
        self.register_buffer("fb_range", filter_ranges(fb), persistent=False)
        self.register_buffer("twiddle", torch.from_numpy(tw.astype(np.float32)).contiguous(), persistent=False)

    @torch.no_grad()
    def forward(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        ops._require_cuda(wave, "log_mel")
        if wave.dim() == 1:
            wave = wave.unsqueeze(0)
        wave = wave.float()
        if wave.stride(-1) != 1:
            wave = wave.contiguous()
        B, N = wave.shape
        n_frames = 1 + N // self.hop_length
        out = torch.empty(B, self.n_mels, n_frames, dtype=torch.float32, device=wave.device)
        lens = None if lengths is None else lengths.to(device=wave.device, dtype=torch.int32)
        ops._run("as_log_mel", wave, wave.data_ptr(), wave.stride(0), ops._p(lens), B, N, self.window.data_ptr(),
                 self.twiddle.data_ptr(), self.fb.data_ptr(), self.fb_range.data_ptr(), self.n_fft, self.hop_length,
                 self.n_mels, float(LOG_EPS), float(self.mean), float(self.std), out.data_ptr(), n_frames)
        return out


_default = {}


def preprocess(wave, device="cuda") -> torch.Tensor:
    """The reference's ``preprocess(wave)`` (test.py:43-47): numpy / tensor ``[N]`` -> ``[1, 80, T]`` on ``device``."""
    dev = torch.device(device)
    if dev not in _default:
        _default[dev] = LogMel().to(dev)
    w = torch.from_numpy(wave) if isinstance(wave, np.ndarray) else wave
    return _default[dev](w.to(dev))
