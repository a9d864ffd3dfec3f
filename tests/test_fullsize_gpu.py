"""Parity at BASELINE.json's full sizes (configs[2] vocoder sweep up to 64 x 3200 frames, configs[3]
long-form 60 s utterances).  The CPU oracle cannot run 128 TFLOP, so the large cases are checked through
size-independent properties of the path — a batch equals its batch-1 runs, a long input equals its own
prefix away from the truncation point — anchored on oracle comparisons of slices the oracle finishes in
seconds, with the north_star tolerances (mel max-abs 1e-2 / mean-abs 1e-3, waveform SNR >= 35 dB)."""
import pytest
import torch

from oracle import restate
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"
RF_FRAMES = 40      # > receptive field of the generator in mel frames (conv_pre 3 + resblocks <= 12 + ups)


def _mel(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 80, T, generator=g).clamp(-2, 2)


def test_vocoder_sweep_max_size_properties():
    gen = util.generator(0).to(DEV)
    B, T = 64, 3200                                            # 204 800 frames, 127.7 TFLOP, 61.4 M samples
    mel = _mel(B, T, 0).to(DEV)
    wav = gen(mel)
    assert wav.shape == (B, 1, 300 * T)
    assert bool(torch.isfinite(wav).all()) and float(wav.abs().max()) <= 1.0
    # (1) a batch is its batch-1 runs, bit for bit (tiles never straddle utterances)
    for i in (0, 37, 63):
        one = gen(mel[i:i + 1])
        assert torch.equal(one[0], wav[i]), f"item {i} differs from its batch-1 run"
    # (2) prefix property: the first frames of a long utterance equal the vocoding of the truncated mel,
    #     except within the receptive field of the truncation point
    for Tp in (200, 800):
        short = gen(mel[5:6, :, :Tp])
        n = (Tp - RF_FRAMES) * 300
        assert torch.equal(short[0, :, :n], wav[5, :, :n]), f"prefix of {Tp} frames differs"
    # (3) anchor: the 200-frame case against the fp32 oracle
    sd = util.generator(0).cpu().state_dict()
    ref = restate.generator_forward(sd, mel[5:6, :, :200].cpu())
    s = util.snr_db(gen(mel[5:6, :, :200]).cpu(), ref)
    assert s >= util.WAV_SNR_DB, f"SNR {s:.1f} dB"
    util._MODELS.clear()


@pytest.mark.parametrize("B,T", [(1, 200), (8, 800), (64, 200), (2, 3200)])
def test_vocoder_sweep_ragged_lengths(B, T):
    """Every sweep point with ragged lengths: padded frames are silent and each item equals its own run."""
    gen = util.generator(0).to(DEV)
    mel = _mel(B, T, 1).to(DEV)
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    wav = gen(mel, lens.to(DEV))
    for i in sorted({0, B - 1}):
        L = int(lens[i])
        assert float(wav[i, :, L * 300:].abs().max()) == 0.0 if L < T else True
        one = gen(mel[i:i + 1, :, :L].contiguous())
        assert torch.equal(one[0, :, :L * 300], wav[i, :, :L * 300]), (B, T, i)
    util._MODELS.clear()


def test_long_form_60s_against_oracle():
    """configs[3]: 900 tokens, sum(dur) = 2400 -> 4800 mel frames = 60 s.  One long utterance next to a
    shorter one in the same batch; the long one is compared with the CPU oracle end to end."""
    g = util.load_golden("acoustic_small.pt")
    model = util.acoustic_model(g["checkpoint_seed"])
    model.set_compute_dtype(torch.float16)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    dist_cpu = {k: v.cpu() for k, v in model.distribution.items()}
    model = model.to(DEV)
    model.distribution = {k: v.to(DEV) for k, v in model.distribution.items()}
    gen = util.generator(0).to(DEV)
    gs = torch.Generator().manual_seed(11)
    Tt, Tr = 900, 240
    tok = torch.randint(1, 178, (2, Tt), generator=gs)
    mel_ref = (torch.randn(2, 80, Tr, generator=gs) * 0.5).clamp(-2, 2)
    dur = torch.full((2, Tt), 2, dtype=torch.long)
    dur[:, :600] += 1                                          # 2 * 900 + 600 = 2400 half-rate frames
    tl = torch.tensor([Tt, 400])
    ml = torch.tensor([Tr, Tr])
    out, aux = model([tok.to(DEV), tl.to(DEV), mel_ref.to(DEV), ml.to(DEV)], step="test", durations=dur.to(DEV),
                     return_aux=True)
    lens_m = aux["mel_lengths"].cpu()
    assert int(lens_m[0]) == 4800 and int(lens_m[1]) == 2 * int(dur[1, :400].sum())
    wav = gen(out, lens_m.to(DEV)).cpu()
    # oracle, utterance 0 (batch-1, fp32, CPU)
    mel_o = restate.artsspeech_test(sd, tok[0:1], mel_ref[0:1], dist_cpu, durations=dur[0])
    d = (out[0:1, :, :4800].cpu() - mel_o).abs()
    assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS, (d.max().item(), d.mean().item())
    wav_o = restate.generator_forward(util.generator(0).cpu().state_dict(), mel_o)
    s = util.snr_db(wav[0:1], wav_o)
    assert s >= util.WAV_SNR_DB, f"SNR {s:.1f} dB"
    # the short utterance: batch-1 run gives the same mel (ragged batch == batch-1 semantics)
    Ls = int(lens_m[1])
    out1, _ = model([tok[1:2, :400].to(DEV), tl[1:2].to(DEV), mel_ref[1:2].to(DEV), ml[1:2].to(DEV)], step="test",
                    durations=dur[1:2, :400].to(DEV), return_aux=True)
    # not bit-exact: the padded length picks other kernel variants (cluster vs TMA InstanceNorm, tile counts),
    # i.e. other fp32 summation orders, and 16-bit operand rounding amplifies the last bits; well inside the
    # mel tolerance
    assert (out1[0, :, :Ls] - out[1, :, :Ls]).abs().max().item() <= 0.5 * util.MEL_MAX_ABS
    assert float(out[1, :, Ls:].abs().max()) == 0.0
    util._MODELS.clear()
