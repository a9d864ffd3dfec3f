"""GPU parity of the CUDA path (through the C ABI) against fixtures generated from the unmodified
reference, with the tolerances BASELINE.json's north_star states:
MAS bit-exact, mel max-abs <= 1e-2 and mean-abs <= 1e-3, waveform SNR >= 35 dB."""
import numpy as np
import pytest
import torch

from artspeech_b200 import mas
from oracle import mas_oracle, restate
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_mas_goldens_bit_exact():
    for key, v, xl, yl, p1, p2 in util.mas_cases():
        vt = torch.from_numpy(v).to(DEV)
        keep = vt.clone()
        xlt, ylt = torch.from_numpy(xl).to(DEV), torch.from_numpy(yl).to(DEV)
        out2 = mas.maximum_path_lens(vt, xlt, ylt, mas.TIE_STAY).cpu().numpy()
        out1 = mas.maximum_path_lens(vt, xlt, ylt, mas.TIE_MOVE).cpu().numpy()
        assert np.array_equal(out2, p2), key
        assert np.array_equal(out1, p1), key
        assert torch.equal(vt, keep), "input must not be modified"


def test_mas_reference_signatures():
    key, v, xl, yl, p1, p2 = next(iter(util.mas_cases()))
    B, Tx, Ty = v.shape
    mask = ((np.arange(Tx)[None, :, None] < xl[:, None, None]) & (np.arange(Ty)[None, None, :] < yl[:, None, None]))
    vt, mt = torch.from_numpy(v).to(DEV), torch.from_numpy(mask.astype(np.float32)).to(DEV)
    assert np.array_equal(mas.maximum_path1(vt, mt).cpu().numpy(), p1)
    assert np.array_equal(mas.maximum_path2(vt, mt).cpu().numpy(), p2)
    assert np.array_equal(mas.maximum_path(vt, mt.bool()).cpu().numpy(), p2)
    m2 = mas.mask_from_lens(vt, torch.from_numpy(xl).to(DEV), torch.from_numpy(yl).to(DEV))
    assert torch.equal(m2, mt)


@pytest.mark.parametrize("kind", ["randn", "ties"])
def test_mas_full_size_bit_exact(kind):
    """BASELINE config 5: 64 x 200 tokens x 1000 frames."""
    rng = np.random.default_rng(7)
    B, Tx, Ty = 64, 200, 1000
    v = rng.standard_normal((B, Tx, Ty)).astype(np.float32) if kind == "randn" else \
        rng.integers(0, 4, (B, Tx, Ty)).astype(np.float32)
    xl = rng.integers(100, 201, B); yl = np.maximum(rng.integers(500, 1001, B), xl)
    xl[0], yl[0] = 200, 1000
    for tie, name in ((mas.TIE_STAY, "stay"), (mas.TIE_MOVE, "move")):
        out = mas.maximum_path_lens(torch.from_numpy(v).to(DEV), torch.from_numpy(xl), torch.from_numpy(yl), tie)
        ref = mas_oracle.maximum_path(v, xl, yl, name)
        assert np.array_equal(out.cpu().numpy(), ref), (kind, name)
        # size-independent properties: one cell per valid column, monotone, ends at the corners
        p = out.cpu().numpy()
        for b in range(0, B, 9):
            cols = p[b, :, :yl[b]].sum(0)
            assert np.all(cols == 1) and p[b, :, yl[b]:].sum() == 0 and p[b, xl[b]:, :].sum() == 0
            rows = p[b, :, :yl[b]].argmax(0)
            assert rows[0] == 0 or xl[b] > yl[b]
            assert rows[-1] == xl[b] - 1 and np.all(np.diff(rows) >= 0) and np.all(np.diff(rows) <= 1)


def test_mas_edge_cases():
    v = torch.randn(3, 7, 9, device=DEV)
    out = mas.maximum_path_lens(v, torch.tensor([0, 1, 7]), torch.tensor([5, 9, 7]), mas.TIE_STAY).cpu()
    assert out[0].sum() == 0                       # empty item
    assert torch.equal(out[1, 0], torch.ones(9)) and out[1, 1:].sum() == 0   # single token
    assert torch.equal(out[2, :, :7], torch.eye(7)) and out[2, :, 7:].sum() == 0   # x_len == y_len: diagonal


@pytest.mark.parametrize("shape", [(5, 37, 90), (8, 200, 1000), (3, 300, 1400)])
def test_mas_align_training_chain(shape):
    """The chain MAS sits in at train_second.py:178-187 / models.py:323-324: softmax -> mask_from_lens ->
    maximum_path -> d_gt = path.sum(-1) and T_en @ path.  as_mas_align returns path, durations and the
    frame -> token map from one launch; as_expand_tokens replaces the dense matmul.  All bit-exact."""
    B, Tx, Ty = shape
    rng = np.random.default_rng(11)
    feat = torch.from_numpy(rng.standard_normal((B, Tx, Ty)).astype(np.float32) * 3.0)
    attn = torch.softmax(feat, dim=-1)
    xl = rng.integers(max(1, Tx // 2), Tx + 1, B); yl = np.maximum(rng.integers(Ty // 2, Ty + 1, B), xl)
    xl[0], yl[0] = Tx, Ty
    xlt, ylt = torch.from_numpy(xl).to(DEV), torch.from_numpy(yl).to(DEV)
    mask = mas.mask_from_lens(attn.to(DEV), xlt, ylt)
    path, dur, tok = mas.align(attn.to(DEV), mask)
    ref = mas_oracle.maximum_path(attn.numpy(), xl, yl, "stay")
    assert np.array_equal(path.cpu().numpy(), ref)
    assert np.array_equal(dur.cpu().numpy(), ref.sum(-1).astype(np.int32))           # d_gt
    tok_ref = np.where(np.arange(Ty)[None, :] < yl[:, None], ref.argmax(1), -1)
    assert np.array_equal(tok.cpu().numpy(), tok_ref)
    T_en = torch.from_numpy(rng.standard_normal((B, 64, Tx)).astype(np.float32))
    want = T_en @ torch.from_numpy(ref)                                               # models.py:323
    got = mas.expand_tokens(T_en.to(DEV), tok).cpu()
    assert torch.equal(got, want)
    # the plain entry point still agrees (shared kernel, null outputs)
    assert torch.equal(mas.maximum_path(attn.to(DEV), mask), path)


def test_vocoder_vs_reference_golden():
    g = util.load_golden("vocoder_small.pt")
    gen = util.generator(g["checkpoint_seed"]).to(DEV)
    gen.compute_dtype = torch.bfloat16
    wav = gen(g["mel"].to(DEV)).cpu()
    assert wav.shape == g["wav"].shape
    s = util.snr_db(wav, g["wav"])
    assert s >= util.WAV_SNR_DB, f"SNR {s:.1f} dB"
    lens = torch.tensor([24, 11])
    wav_r = gen(g["mel"].to(DEV), lens.to(DEV)).cpu()
    one = restate.generator_forward(util.generator(g["checkpoint_seed"]).cpu().state_dict(), g["mel"][1:2, :, :11])
    assert util.snr_db(wav_r[1:2, :, :11 * 300], one) >= util.WAV_SNR_DB
    assert wav_r[1, :, 11 * 300:].abs().max().item() == 0.0
    # 16-bit PCM straight from the conv_post kernel == what soundfile.write stores for the float waveform
    # (test.py:119; libsndfile: lrint(32767 * x)), and still >= 35 dB against the reference waveform
    pcm = gen(g["mel"].to(DEV), pcm16=True).cpu()
    assert pcm.dtype == torch.int16 and pcm.shape == wav.shape
    assert torch.equal(pcm, torch.round(wav * 32767.0).clamp(-32768, 32767).to(torch.int16))
    assert util.snr_db(pcm.float() / 32767.0, g["wav"]) >= util.WAV_SNR_DB
    # chunked vocoding with a receptive-field halo reproduces the one-shot pass
    mel = g["mel"].to(DEV)
    parts = list(gen.stream(mel, chunk_frames=8))
    assert [p[0] for p in parts] == [i * 8 * 300 for i in range(len(parts))]
    cat = torch.cat([p[1] for p in parts], dim=2).cpu()
    assert cat.shape == wav.shape and util.snr_db(cat, wav) >= 60.0
    with pytest.raises(Exception):      # PCM is only defined for the single-channel output convolution
        from artspeech_b200 import ops
        ops.conv(torch.zeros(1, 128, 64, device=DEV, dtype=torch.bfloat16), gen._plan["ups"][3], act_out=torch.int16)
    util._MODELS.clear()


@pytest.fixture(scope="module")
def model_gpu():
    g = util.load_golden("acoustic_small.pt")
    m = util.acoustic_model(g["checkpoint_seed"])
    m.set_compute_dtype(torch.float16)
    m = m.to(DEV)
    m.distribution = {k: v.to(DEV) for k, v in m.distribution.items()}
    yield m, g
    util._MODELS.clear()


@pytest.mark.parametrize("case", ["a_pred_dur", "b_forced_dur", "c_short"])
def test_acoustic_vs_reference_golden(model_gpu, case):
    model, g = model_gpu
    c = g["cases"][case]
    tok, mel = c["tokens"].to(DEV), c["ref_mel"].to(DEV)
    dur = c["out"]["pred_dur"].view(1, -1).to(DEV)      # durations fed from the reference's integer output
    out, aux = model([tok, torch.tensor([tok.shape[1]], device=DEV), mel, torch.tensor([mel.shape[2]], device=DEV)],
                     step="test", durations=dur, return_aux=True)
    o = c["out"]
    d = (out.cpu() - o["mel"]).abs()
    assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS, (d.max().item(), d.mean().item())
    assert (aux["style"].cpu() - o["style"]).abs().max().item() < 2e-3
    for k in ("F0", "N", "EMA"):
        assert (aux[k].transpose(1, 2).cpu() - o[k]).abs().max().item() < 2e-3, k
    for k in ("f0_ext", "n_ext", "ema_ext"):
        assert (aux[k].cpu() - o[k]).abs().max().item() < 2e-3, k
    if case == "a_pred_dur":   # the duration predictor itself (rounding must agree with the reference)
        out2, aux2 = model([tok, torch.tensor([tok.shape[1]], device=DEV), mel, torch.tensor([mel.shape[2]], device=DEV)],
                           step="test", return_aux=True)
        assert (aux2["duration"].cpu() - o["duration"]).abs().max().item() < 2e-3
        assert torch.equal(aux2["pred_dur"].long().view(-1).cpu(), o["pred_dur"].view(-1))


def test_ragged_batch_and_text_to_wave(model_gpu):
    model, g = model_gpu
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
    out, aux = model([tok.to(DEV), tl.to(DEV), mel.to(DEV), ml.to(DEV)], step="test", durations=dur.to(DEV),
                     return_aux=True)
    lens_m = aux["mel_lengths"].cpu()
    for i, c in enumerate(cs):
        Tm = c["out"]["mel"].shape[2]
        assert int(lens_m[i]) == Tm
        d = (out[i, :, :Tm].cpu() - c["out"]["mel"][0]).abs()
        assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS, (names[i], d.max().item())
    # text -> waveform: vocode our mel, compare with the fp32 oracle vocoding the reference mel
    gen = util.generator(0).to(DEV)
    wav = gen(out, lens_m.to(DEV)).cpu()
    sd = util.generator(0).cpu().state_dict()
    for i, c in enumerate(cs):
        Tm = c["out"]["mel"].shape[2]
        ref_wav = restate.generator_forward(sd, c["out"]["mel"])
        s = util.snr_db(wav[i:i + 1, :, :Tm * 300], ref_wav)
        assert s >= util.WAV_SNR_DB, f"{names[i]}: SNR {s:.1f} dB"


def test_mixed_length_micro_batches_are_batch_invariant(model_gpu):
    """BASELINE config 5 in small: mixed-length utterances through engine.synthesize_many (length-bucketed ragged
    micro-batches) give every utterance the waveform of its own batch-1 run, whatever it was batched with."""
    from artspeech_b200 import engine
    model, g = model_gpu
    gen = util.generator(0).to(DEV)
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=False)
    gsrc = torch.Generator().manual_seed(21)
    tl = [15, 90, 33, 61, 47, 120, 18, 75, 52]
    toks = [torch.randint(1, 178, (t,), generator=gsrc) for t in tl]
    durs = [torch.randint(1, 4, (t,), generator=gsrc) for t in tl]
    mels = [(torch.randn(80, 120, generator=gsrc) * 0.5).clamp(-2, 2) for _ in tl]
    wavs, frames = engine.synthesize_many(syn, toks, mels, durs, max_batch=4, max_padded_frames=1200)
    assert frames == [2 * int(d.sum()) for d in durs]
    for i in (0, 1, 5, 8):
        one, fr1 = engine.synthesize_many(syn, [toks[i]], [mels[i]], [durs[i]])
        assert wavs[i].numel() == 300 * frames[i] == one[0].numel()
        s = util.snr_db(wavs[i].cpu().view(1, 1, -1), one[0].cpu().view(1, 1, -1))
        assert s >= 50.0, f"utterance {i}: {s:.1f} dB vs its batch-1 run"


def test_pipelined_engine_matches_serial(model_gpu):
    """Two batches in flight (engine.Synthesizer(pipeline_depth=2): CUDA-graph slots on side streams) give
    exactly the waveforms of the serial engine, call after call, with inputs changing between calls."""
    from artspeech_b200 import engine
    model, g = model_gpu
    gen = util.generator(0).to(DEV)
    B, Tt, Tr = 3, 40, 120
    gsrc = torch.Generator().manual_seed(5)
    tl, ml = torch.full((B,), Tt), torch.full((B,), Tr)
    dur = torch.randint(1, 4, (B, Tt), generator=gsrc)
    serial = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, pipeline_depth=1)
    piped = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, pipeline_depth=2)
    hosts = []
    want = []
    for step in range(5):
        tok = torch.randint(1, 178, (B, Tt), generator=gsrc).to(DEV)
        mel = (torch.randn(B, 80, Tr, generator=gsrc) * 0.5).to(DEV)
        w_ref, _, _ = serial.synthesize(tok, tl, mel, ml, dur)
        want.append(w_ref.clone().cpu())
        w, _, _ = piped.synthesize(tok, tl, mel, ml, dur)
        h = torch.empty(w.shape, dtype=w.dtype).pin_memory()
        with torch.cuda.stream(piped.last_stream):
            h.copy_(w, non_blocking=True)
        hosts.append(h)
    piped.join()
    torch.cuda.synchronize()
    for step, (h, w) in enumerate(zip(hosts, want)):
        assert torch.equal(h, w), f"step {step}: pipelined output differs"


def test_voice_cache_matches_full_pass(model_gpu):
    """engine.Synthesizer.encode_voice + synthesize(voice=...) (style encoder once per voice)
    gives exactly the waveforms of the full pass, eagerly and through the CUDA-graph path."""
    from artspeech_b200 import engine
    model, g = model_gpu
    gen = util.generator(0).to(DEV)
    B, Tt, Tr = 2, 30, 100
    gsrc = torch.Generator().manual_seed(9)
    tok = torch.randint(1, 178, (B, Tt), generator=gsrc).to(DEV)
    mel = (torch.randn(B, 80, Tr, generator=gsrc) * 0.5).to(DEV)
    tl, ml = torch.full((B,), Tt), torch.full((B,), Tr)
    dur = torch.randint(1, 4, (B, Tt), generator=gsrc)
    for use_graph in (False, True):
        syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=use_graph)
        full, _, mel_full = syn.synthesize(tok, tl, mel, ml, dur)
        full, mel_full = full.clone(), mel_full.clone()
        voice = syn.encode_voice(mel, ml)
        for _ in range(2):
            cached, _, mel_cached = syn.synthesize(tok, tl, mel, ml, dur, voice=voice)
        assert torch.equal(mel_cached, mel_full) and torch.equal(cached, full), f"graph={use_graph}"
