"""CPU tests (no GPU): host logic of the module classes on the torch semantic backend, and the MAS
oracle, against fixtures generated from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import mas_oracle, restate
from tests import util


def test_mas_oracle_matches_reference_goldens():
    n = 0
    for key, v, xl, yl, p1, p2 in util.mas_cases():
        assert np.array_equal(mas_oracle.maximum_path(v, xl, yl, "move"), p1), key
        assert np.array_equal(mas_oracle.maximum_path(v, xl, yl, "stay"), p2), key
        n += 1
    assert n == 6


def test_mas_tie_modes_differ_on_ties():
    diff = 0
    for key, v, xl, yl, p1, p2 in util.mas_cases():
        if "randn" in key:
            assert np.array_equal(p1, p2)
        else:
            diff += int(not np.array_equal(p1, p2))
    assert diff >= 1   # v1 and v2 disagree on tie-heavy inputs


def test_vocoder_receptive_field_bounds_the_reference():
    """Generator.receptive_field_frames() (the halo of the chunked `stream`) checked on the reference restatement:
    a chunk vocoded with that much real context on both sides reproduces the full pass in its centre; one frame of
    context does not."""
    gen = util.generator(0)
    sd = gen.state_dict()
    torch.manual_seed(5)
    mel = torch.randn(1, 80, 72).clamp(-2, 2)
    full = restate.generator_forward(sd, mel)
    halo = gen.receptive_field_frames()
    assert 10 <= halo <= 16
    t0, t1 = 30, 40
    part = restate.generator_forward(sd, mel[:, :, t0 - halo:t1 + halo])[:, :, halo * 300:(halo + t1 - t0) * 300]
    assert (part - full[:, :, t0 * 300:t1 * 300]).abs().max().item() < 1e-6
    short = restate.generator_forward(sd, mel[:, :, t0 - 1:t1 + 1])[:, :, 300:(1 + t1 - t0) * 300]
    assert (short - full[:, :, t0 * 300:t1 * 300]).abs().max().item() > 1e-4


def test_vocoder_restatement_and_host_logic(sim):
    g = util.load_golden("vocoder_small.pt")
    gen = util.generator(g["checkpoint_seed"])
    ref = restate.generator_forward(gen.state_dict(), g["mel"])
    assert (ref - g["wav"]).abs().max().item() < 1e-5          # oracle pinned to the reference output
    gen.compute_dtype = torch.bfloat16
    gen.invalidate_plan()
    wav = gen(g["mel"])
    assert wav.shape == g["wav"].shape
    assert util.snr_db(wav, g["wav"]) >= util.WAV_SNR_DB
    # ragged batch == batch-1 runs
    lens = torch.tensor([24, 11])
    wav_r = gen(g["mel"], lens)
    one = restate.generator_forward(gen.state_dict(), g["mel"][1:2, :, :11])
    assert util.snr_db(wav_r[1:2, :, :11 * 300], one) >= util.WAV_SNR_DB
    assert wav_r[1, :, 11 * 300:].abs().max().item() == 0.0


@pytest.mark.parametrize("case", ["a_pred_dur", "b_forced_dur", "c_short"])
def test_acoustic_host_logic_fp32_exact(sim, monkeypatch, case):
    """With fp32 'operands' the dataflow must reproduce the reference to rounding noise: this pins
    weight folding, layouts, tap tables, fused concats, masks and the batched style FCs."""
    g = util.load_golden("acoustic_small.pt")
    model = util.acoustic_model(g["checkpoint_seed"])
    model.set_compute_dtype(torch.float32)
    monkeypatch.setattr(sim, "EMULATE_FP16_RECURRENCE", False)
    c = g["cases"][case]
    tok, mel = c["tokens"], c["ref_mel"]
    out, aux = model([tok, torch.tensor([tok.shape[1]]), mel, torch.tensor([mel.shape[2]])], step="test",
                     durations=c["durations"], return_aux=True)
    o = c["out"]
    assert (aux["style"] - o["style"]).abs().max().item() < 1e-5
    assert (aux["T_en"][:, :, :8] - o["T_en"]).abs().max().item() < 1e-4
    assert abs(float(aux["T_en"].double().sum()) - float(o["T_en_sum"])) < 1e-2
    if c["durations"] is None:
        assert (aux["duration"] - o["duration"]).abs().max().item() < 1e-5
        assert torch.equal(aux["pred_dur"].long().view(-1), o["pred_dur"].view(-1))
    for k in ("F0", "N", "EMA"):
        assert (aux[k].transpose(1, 2) - o[k]).abs().max().item() < 1e-5, k
    assert out.shape == o["mel"].shape
    assert (out - o["mel"]).abs().max().item() < 1e-4


def test_acoustic_fp16_within_tolerance(sim):
    g = util.load_golden("acoustic_small.pt")
    model = util.acoustic_model(g["checkpoint_seed"])
    model.set_compute_dtype(torch.float16)
    c = g["cases"]["b_forced_dur"]
    tok, mel = c["tokens"], c["ref_mel"]
    out = model([tok, torch.tensor([tok.shape[1]]), mel, torch.tensor([mel.shape[2]])], step="test",
                durations=c["durations"])
    d = (out - c["out"]["mel"]).abs()
    assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS


def test_ragged_batch_equals_batch1(sim, monkeypatch):
    """B=3 with different token / reference lengths must equal three batch-1 reference runs."""
    g = util.load_golden("acoustic_small.pt")
    model = util.acoustic_model(g["checkpoint_seed"])
    model.set_compute_dtype(torch.float32)
    monkeypatch.setattr(sim, "EMULATE_FP16_RECURRENCE", False)
    names = ["a_pred_dur", "b_forced_dur", "c_short"]
    cs = [g["cases"][n] for n in names]
    Tt = max(c["tokens"].shape[1] for c in cs)
    Tr = max(c["ref_mel"].shape[2] for c in cs)
    tok = torch.zeros(3, Tt, dtype=torch.long)
    mel = torch.zeros(3, 80, Tr)
    dur = torch.ones(3, Tt, dtype=torch.long)
    for i, c in enumerate(cs):
        tok[i, :c["tokens"].shape[1]] = c["tokens"][0]
        mel[i, :, :c["ref_mel"].shape[2]] = c["ref_mel"][0]
        dur[i, :c["tokens"].shape[1]] = c["out"]["pred_dur"].view(-1)
    tl = torch.tensor([c["tokens"].shape[1] for c in cs])
    ml = torch.tensor([c["ref_mel"].shape[2] for c in cs])
    out, aux = model([tok, tl, mel, ml], step="test", durations=dur, return_aux=True)
    for i, c in enumerate(cs):
        Tm = c["out"]["mel"].shape[2]
        assert int(aux["mel_lengths"][i]) == Tm
        assert (out[i, :, :Tm] - c["out"]["mel"][0]).abs().max().item() < 1e-4, names[i]
        if Tm < out.shape[2]:
            assert out[i, :, Tm:].abs().max().item() == 0.0
