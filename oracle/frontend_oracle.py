"""TEST INFRASTRUCTURE — CPU restatement of the reference's log-mel front-end.

Reference: ``test.py:40-47`` / ``meldataset.py:42-49``::

    to_mel = torchaudio.transforms.MelSpectrogram(n_mels=80, n_fft=2048, win_length=1200, hop_length=300)
    mel = (torch.log(1e-5 + to_mel(wave).unsqueeze(0)) - (-4)) / 4

The arithmetic lives in a third-party dependency, torchaudio (unpinned by the reference; 2.11.0 in this
image): ``MelSpectrogram`` = ``Spectrogram`` (torch.stft, centre = True with reflect padding of n_fft/2,
periodic Hann window of win_length zero-padded symmetrically to n_fft, one-sided, power 2, not normalised)
followed by ``MelScale`` (HTK mel scale, triangular filters without area normalisation, f_min 0,
f_max = sample_rate / 2 with the class default sample_rate = 16000 — the reference never passes its
24 kHz rate, so the filterbank spans "8 kHz" of a 16 kHz axis; this is reproduced, not fixed).
Pinned by ``tests/golden/frontend.pt`` (generated with the real torchaudio by ``oracle/make_golden_frontend.py``).
Only tests / smoke / bench's CPU leg may import this module.
"""
from __future__ import annotations

import math

import torch

N_FFT, WIN, HOP, N_MELS = 2048, 1200, 300, 80
SAMPLE_RATE_DEFAULT = 16000       # torchaudio's default, what the reference effectively uses
LOG_EPS, MEAN, STD = 1e-5, -4.0, 4.0


def hann_window_padded() -> torch.Tensor:
    """Periodic Hann window of WIN samples, zero-padded symmetrically to N_FFT (torch.stft semantics)."""
    w = torch.hann_window(WIN, periodic=True, dtype=torch.float64)
    left = (N_FFT - WIN) // 2
    out = torch.zeros(N_FFT, dtype=torch.float64)
    out[left:left + WIN] = w
    return out.float()


def mel_filterbank() -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(1025, 0, 8000, 80, 16000, norm=None, mel_scale='htk') -> [1025, 80]."""
    n_freqs = N_FFT // 2 + 1
    all_freqs = torch.linspace(0, SAMPLE_RATE_DEFAULT // 2, n_freqs)
    hz_to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    m_pts = torch.linspace(hz_to_mel(0.0), hz_to_mel(SAMPLE_RATE_DEFAULT / 2.0), N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)          # [n_freqs, n_mels + 2]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def log_mel(wave: torch.Tensor) -> torch.Tensor:
    """``wave`` fp32 [N] or [B, N] -> normalised log-mel [B, 80, 1 + N // 300] (fp32)."""
    if wave.dim() == 1:
        wave = wave.unsqueeze(0)
    wave = wave.float()
    pad = N_FFT // 2
    x = torch.nn.functional.pad(wave.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(1, N_FFT, HOP)                                # [B, n_frames, N_FFT]
    spec = torch.fft.rfft(frames * hann_window_padded(), dim=-1)    # one-sided DFT
    power = spec.real ** 2 + spec.imag ** 2                         # [B, n_frames, 1025]
    mel = power @ mel_filterbank()                                  # [B, n_frames, 80]
    return ((torch.log(LOG_EPS + mel) - MEAN) / STD).transpose(1, 2).contiguous()
