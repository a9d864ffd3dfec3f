"""TEST INFRASTRUCTURE — writes tests/golden/frontend.pt with the REAL torchaudio transform the reference
builds (test.py:40-47, meldataset.py:42-49):  python -m oracle.make_golden_frontend"""
import os

import torch
import torchaudio

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frontend.pt")

# --- the reference's own lines (test.py:40-47) ---
to_mel = torchaudio.transforms.MelSpectrogram(n_mels=80, n_fft=2048, win_length=1200, hop_length=300)
mean, std = -4, 4


def preprocess(wave_tensor):
    mel_tensor = to_mel(wave_tensor)
    mel_tensor = (torch.log(1e-5 + mel_tensor.unsqueeze(0)) - mean) / std
    return mel_tensor
# -------------------------------------------------


def synth_wave(n, seed):
    """Speech-like test signal: harmonics of a gliding pitch with a formant-ish envelope, plus noise and a
    silent stretch (exercises the 1e-5 floor)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64) / 24000.0
    f0 = 110.0 + 40.0 * torch.sin(2 * torch.pi * 0.7 * t)
    phase = 2 * torch.pi * torch.cumsum(f0, 0) / 24000.0
    x = sum((0.5 / h) * torch.sin(h * phase) for h in range(1, 30))
    x = x * (0.3 + 0.2 * torch.sin(2 * torch.pi * 3.1 * t)) + 0.01 * torch.randn(n, generator=g, dtype=torch.float64)
    x[n // 3: n // 3 + 2500] = 0.0
    return (0.4 * x / x.abs().max()).float()


def main():
    cases = {}
    for name, n, seed in (("a_1s", 24000, 1), ("b_3s", 72000, 2), ("c_odd", 30011, 3), ("d_short", 2500, 4)):
        w = synth_wave(n, seed)
        cases[name] = {"wave": w, "mel": preprocess(w).squeeze(0).clone()}      # [80, 1 + n // 300]
    torch.save({"cases": cases, "fb": to_mel.mel_scale.fb.clone(), "window": to_mel.spectrogram.window.clone(),
                "torchaudio": torchaudio.__version__}, OUT)
    print("wrote", OUT, {k: tuple(v["mel"].shape) for k, v in cases.items()})


if __name__ == "__main__":
    main()
