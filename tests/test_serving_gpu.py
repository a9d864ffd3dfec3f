"""The serving engine on the GPU: bucketed CUDA graphs with tokens / reference mels / lengths / DURATIONS as graph
inputs, the duration predictor on the device (models.py:360-361), the LRU graph cache, and parity at the
benchmark's shape against the CPU oracle (north_star tolerances: mel max-abs 1e-2 / mean-abs 1e-3, SNR >= 35 dB)."""
import pytest
import torch

from oracle import restate
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def model_gpu():
    g = util.load_golden("acoustic_small.pt")
    m = util.acoustic_model(g["checkpoint_seed"])
    m.set_compute_dtype(torch.float16)
    sd = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    dist_cpu = {k: v.cpu().clone() for k, v in m.distribution.items()}
    m = m.to(DEV)
    m.distribution = {k: v.to(DEV) for k, v in m.distribution.items()}
    yield m, g, sd, dist_cpu
    util._MODELS.clear()


def test_round_durations_kernel():
    """as_round_durations == torch.round(x).clamp(min=1) (round-half-to-even), zero beyond the length, with sums."""
    from artspeech_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, Tt = 7, 93
    pred = torch.randn(B, Tt, generator=g) * 3 + 2
    pred[0, :8] = torch.tensor([0.5, 1.5, 2.5, 3.5, -0.5, 0.49999, 1.0, -7.0])        # ties, negatives
    lens = torch.tensor([93, 1, 50, 92, 33, 64, 2], dtype=torch.int32)
    dur, sums = ops.round_durations(pred.to(DEV), lens.to(DEV))
    want = torch.round(pred).clamp(min=1).to(torch.int32)
    want = want * (torch.arange(Tt)[None, :] < lens[:, None]).to(torch.int32)
    assert torch.equal(dur.cpu(), want) and torch.equal(sums.cpu(), want.sum(1).to(torch.int32))
    # strided rows (a [B,Tt] view of the projection's [B,Tt,1] output is contiguous; a column slice is not)
    wide = torch.randn(B, Tt + 5, generator=g).to(DEV)
    dur2, _ = ops.round_durations(wide[:, :Tt], None)
    assert torch.equal(dur2.cpu(), torch.round(wide[:, :Tt].cpu()).clamp(min=1).to(torch.int32))


def test_conv_emits_a_second_16bit_format():
    """The decoder's last conv writes the fp32 mel AND its bf16 copy for the vocoder from fp16 operands."""
    from artspeech_b200 import nn_util, ops
    g = torch.Generator().manual_seed(1)
    conv = torch.nn.Conv1d(512, 80, 1)
    x = torch.randn(3, 130, 512, generator=g)
    lens = torch.tensor([130, 7, 64], dtype=torch.int32, device=DEV)
    pw = nn_util.pack_conv1d(conv, torch.float16, torch.device(DEV))
    raw, b16 = ops.conv(x.to(DEV).half(), pw, raw=torch.float32, act_out=torch.bfloat16, lens=lens)
    ref = torch.nn.functional.conv1d(x.half().float().transpose(1, 2), conv.weight.detach().half().float(),
                                     conv.bias.detach()).transpose(1, 2)
    keep = (torch.arange(130)[None, :] < lens.cpu()[:, None]).unsqueeze(-1)
    assert (raw.cpu() - ref * keep).abs().max().item() < 2e-3
    assert torch.equal(b16.cpu(), raw.cpu().to(torch.bfloat16))


def _batch(gsrc, B, Tt, Tr, tl):
    tok = torch.randint(1, 178, (B, Tt), generator=gsrc)
    mel = (torch.randn(B, 80, Tr, generator=gsrc) * 0.5).clamp(-2, 2)
    dur = torch.randint(1, 4, (B, Tt), generator=gsrc)
    for b in range(B):
        tok[b, tl[b]:] = 0
    return tok, mel, dur


def test_graph_inputs_change_every_call(model_gpu):
    """ONE captured graph serves calls whose tokens, reference mels, token lengths and durations all differ (same
    shape bucket); every call equals the eager pass on the same inputs, and each utterance's frame count follows
    its own durations."""
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    graph = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True)
    eager = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=False)
    gsrc = torch.Generator().manual_seed(3)
    B, Tr = 3, 120
    for step, (Tt, tl) in enumerate([(40, [40, 33, 12]), (37, [20, 37, 36]), (33, [33, 33, 5]), (40, [1, 40, 17])]):
        tok, mel, dur = _batch(gsrc, B, Tt, Tr, tl)
        dur[tl.index(max(tl))] = 3                                # longest utterance: 99..120 half-rate frames, so
        sums = [int(dur[b, :tl[b]].sum()) for b in range(B)]      # every call lands in the (80, 120] frame bucket
        assert 80 < max(sums) <= 120
        wg, lg, mg = graph.synthesize(tok.to(DEV), tl, mel.to(DEV), [Tr] * B, dur)
        wg, lg, mg = wg.clone(), lg.clone(), mg.clone()
        we, le, me = eager.synthesize(tok.to(DEV), tl, mel.to(DEV), [Tr] * B, dur)
        assert lg.tolist() == le.tolist() == [2 * s for s in sums]
        assert graph.last["frames"] == [2 * s for s in sums]
        for b in range(B):
            n = 2 * sums[b]
            d = (mg[b, :, :n] - me[b, :, :n]).abs()
            assert d.max().item() <= 2e-3, (step, b, d.max().item())
            assert util.snr_db(wg[b, :300 * n].cpu(), we[b, :300 * n].cpu()) >= 50.0, (step, b)
            assert float(wg[b, 300 * n:].abs().max()) == 0.0 if 300 * n < wg.shape[1] else True
    assert graph.stats["captures"] == 1 and graph.stats["replays"] == 4 and graph.stats["eager"] == 0


def test_predicted_durations_on_the_device(model_gpu):
    """durations=None: graph A (encoders + duration predictor + round/clamp on the device) -> B frame counts read
    back -> graph B.  The integer durations equal the reference's (golden case a) and the waveform equals the pass
    that is fed those integers."""
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    c = g["cases"]["a_pred_dur"]
    tok, mel = c["tokens"], c["ref_mel"]
    Tt, Tr = tok.shape[1], mel.shape[2]
    want = c["out"]["pred_dur"].view(-1).long()
    for use_graph in (True, False):
        syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=use_graph)
        for _ in range(2):
            w, lens, m = syn.synthesize(tok.to(DEV), [Tt], mel.to(DEV), [Tr], None)
        assert torch.equal(syn.last["pred_dur"][0, :Tt].cpu().long(), want), f"graph={use_graph}"
        assert syn.last["frames"] == [2 * int(want.sum())] and lens.tolist() == [2 * int(want.sum())]
        d = (m.cpu() - c["out"]["mel"]).abs()
        assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS
        w2, _, _ = syn.synthesize(tok.to(DEV), [Tt], mel.to(DEV), [Tr], want.view(1, -1))
        assert util.snr_db(w.cpu(), w2.cpu()) >= 50.0
        if use_graph:
            assert syn.stats["captures"] == 3 and syn.stats["eager"] == 0      # A, B and the forced-duration graph
    # forced durations with the predictor inside the same graph (the benchmark's step): same waveform, and the
    # predictor's output is still available
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True)
    w3, _, _ = syn.synthesize(tok.to(DEV), [Tt], mel.to(DEV), [Tr], want.view(1, -1), predict_durations=True)
    assert util.snr_db(w3.cpu(), w2.cpu()) >= 50.0
    assert torch.equal(syn.last["pred_dur"][0, :Tt].cpu().long(), want)
    assert (syn.last["duration"][0, :Tt].cpu() - c["out"]["duration"].view(-1)).abs().max().item() < 2e-3


def test_graph_cache_is_bounded_lru(model_gpu):
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, max_graphs=2, capture_after=2)
    ref = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=False)
    gsrc = torch.Generator().manual_seed(8)
    Tr = 100
    shapes = [(2, 20), (2, 50), (1, 70)]                      # three different buckets (token quantum 32)
    data = [_batch(gsrc, B, Tt, Tr, [Tt] * B) for B, Tt in shapes]

    def run(i):
        tok, mel, dur = data[i]
        B, Tt = shapes[i]
        w, _, _ = syn.synthesize(tok.to(DEV), [Tt] * B, mel.to(DEV), [Tr] * B, dur)
        w = w.clone()
        w0, _, _ = ref.synthesize(tok.to(DEV), [Tt] * B, mel.to(DEV), [Tr] * B, dur)
        assert util.snr_db(w.cpu(), w0.cpu()) >= 50.0, i
    run(0)
    assert syn.stats == {"captures": 0, "replays": 0, "eager": 1, "evictions": 0, "drops": 0}   # first sighting: eager
    run(0); run(1); run(1); run(2); run(2)
    assert syn.stats["captures"] == 3 and syn.stats["evictions"] == 1
    assert len(syn._graphs[0]) == 2
    run(0)                                                    # evicted bucket: captured again, still correct
    assert syn.stats["captures"] == 4 and syn.stats["evictions"] == 2
    run(2)
    assert syn.stats["captures"] == 4


def test_synthesize_many_to_host_and_predicted(model_gpu):
    """engine.synthesize_many with pinned-host output and with predicted durations: every utterance equals its
    own batch-1 pass."""
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, pipeline_depth=2)
    gsrc = torch.Generator().manual_seed(21)
    tl = [15, 90, 33, 61, 47, 120, 18]
    toks = [torch.randint(1, 178, (t,), generator=gsrc) for t in tl]
    durs = [torch.randint(1, 4, (t,), generator=gsrc) for t in tl]
    mels = [(torch.randn(80, 120, generator=gsrc) * 0.5).clamp(-2, 2) for _ in tl]
    wavs, frames = engine.synthesize_many(syn, toks, mels, durs, max_batch=3, max_padded_frames=1500, to_host=True)
    assert frames == [2 * int(d.sum()) for d in durs] and all(not w.is_cuda for w in wavs)
    one = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=False)
    for i in (0, 1, 5):
        w1, _, _ = one.synthesize(toks[i].view(1, -1).to(DEV), [tl[i]], mels[i].unsqueeze(0).to(DEV), [120], durs[i].view(1, -1))
        assert wavs[i].numel() == 300 * frames[i]
        assert util.snr_db(wavs[i], w1[0, :300 * frames[i]].cpu()) >= 50.0, i
    # predicted durations: frame counts come back from the device
    wavs_p, frames_p = engine.synthesize_many(syn, toks, mels, None, max_batch=3, to_host=True)
    for i in (2, 6):
        w1, _, _ = one.synthesize(toks[i].view(1, -1).to(DEV), [tl[i]], mels[i].unsqueeze(0).to(DEV), [120], None)
        assert frames_p[i] == one.last["frames"][0] and wavs_p[i].numel() == 300 * frames_p[i]
        assert util.snr_db(wavs_p[i], w1[0, :300 * frames_p[i]].cpu()) >= 50.0, i


def test_bench_shape_against_oracle(model_gpu):
    """The benchmark's own step (16 utterances x 150 tokens, 240-frame reference mels, seeded durations summing to
    400 -> 800 frames, duration predictor inside the graph) against the CPU oracle on one of the 16 utterances:
    predicted integer durations, mel and waveform."""
    import bench
    from artspeech_b200 import engine
    model, g, sd, dist_cpu = model_gpu
    gen = util.generator(0).to(DEV)
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, pipeline_depth=2)
    tok, tl, mel, ml, dur = bench.make_inputs(0, 16, 3)
    assert bool((dur.sum(1) == 400).all())
    for _ in range(2):
        wav, lens, m = syn.synthesize(tok.pin_memory(), tl, mel.pin_memory(), ml, dur, predict_durations=True)
    syn.join()
    torch.cuda.synchronize()
    assert lens.tolist() == [800] * 16 and wav.shape == (16, 240000)
    i = 11
    mel_o, aux = restate.artsspeech_test(sd, tok[i:i + 1], mel[i:i + 1], dist_cpu, durations=dur[i], want_aux=True,
                                         predict_durations=True)
    d = (m[i:i + 1].cpu() - mel_o).abs()
    assert d.max().item() <= util.MEL_MAX_ABS and d.mean().item() <= util.MEL_MEAN_ABS, (d.max().item(), d.mean().item())
    wav_o = restate.generator_forward(util.generator(0).cpu().state_dict(), mel_o)
    s = util.snr_db(wav[i].cpu(), wav_o)
    assert s >= util.WAV_SNR_DB, f"SNR {s:.1f} dB"
    # the duration predictor that ran inside the graph: same integers as the oracle's
    # (|duration - k - 0.5| < 2e-3 would be a legitimate rounding flip; none occurs for this seed)
    want = torch.round(aux["duration"].view(-1)).clamp(min=1).long()
    got = syn.last["pred_dur"][i, :150].cpu().long()
    near_tie = ((aux["duration"].view(-1) % 1.0) - 0.5).abs() < 2e-3
    assert torch.equal(got[~near_tie], want[~near_tie])
    assert (syn.last["duration"][i, :150].cpu() - aux["duration"].view(-1)).abs().max().item() < 5e-3


def test_graphs_survive_a_second_engine_and_follow_weight_reloads(model_gpu):
    """Captured graphs point at the packed weights.  Building another Synthesizer on the same modules
    (``model.to(device)`` on a model that is already there) must not re-pack them; a ``load_state_dict`` does, and
    the engine then re-captures instead of replaying graphs that read freed memory."""
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    gsrc = torch.Generator().manual_seed(13)
    tok, mel, dur = _batch(gsrc, 2, 30, 100, [30, 30])
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True)
    w0, _, _ = syn.synthesize(tok.to(DEV), [30, 30], mel.to(DEV), [100, 100], dur)
    w0 = w0.clone()
    other = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True)      # same modules, same device
    other.synthesize(tok.to(DEV), [30, 30], mel.to(DEV), [100, 100], dur)
    del other
    torch.cuda.empty_cache()
    w1, _, _ = syn.synthesize(tok.to(DEV), [30, 30], mel.to(DEV), [100, 100], dur)
    assert torch.equal(w1, w0) and syn.stats["captures"] == 1
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(sd)                                                    # re-packs the weights
    torch.cuda.empty_cache()
    w2, _, _ = syn.synthesize(tok.to(DEV), [30, 30], mel.to(DEV), [100, 100], dur)
    torch.cuda.synchronize()
    assert torch.equal(w2, w0) and syn.stats["captures"] == 2 and syn.stats["drops"] == 1


def test_from_cache_is_bit_equal_to_fold_on_load(model_gpu, tmp_path):
    """a Synthesizer started from the converted weight cache produces exactly the waveforms of
    one that folds / packs the parameters itself; a wrong content hash is refused."""
    from artspeech_b200 import convert, engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    cache = convert.build_cache(model, gen, include_state=False)
    path = str(tmp_path / "cache.pt")
    torch.save(cache, path)
    with pytest.raises(ValueError):
        engine.Synthesizer.from_cache(path, device=DEV, expect_hash="0" * 64)
    cached = engine.Synthesizer.from_cache(path, device=DEV, expect_hash=cache["hash"], use_cuda_graph=False)
    folded = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=False)
    gsrc = torch.Generator().manual_seed(17)
    tok, mel, dur = _batch(gsrc, 2, 35, 110, [35, 20])
    w0, l0, m0 = folded.synthesize(tok.to(DEV), [35, 20], mel.to(DEV), [110, 110], dur)
    w1, l1, m1 = cached.synthesize(tok.to(DEV), [35, 20], mel.to(DEV), [110, 110], dur)
    assert torch.equal(w0, w1) and torch.equal(m0, m1) and l0.tolist() == l1.tolist()
    # predicted durations through the cached weights as well (duration predictor plan)
    w2, _, _ = folded.synthesize(tok.to(DEV), [35, 20], mel.to(DEV), [110, 110], None)
    w3, _, _ = cached.synthesize(tok.to(DEV), [35, 20], mel.to(DEV), [110, 110], None)
    assert torch.equal(w2, w3)


def test_bounded_single_graph_predicted_path(model_gpu):
    """Synthesizer(duration_bound=...): predicted durations in ONE graph with a generous frame bucket (no host
    synchronisation; the kernels skip the tiles beyond each utterance's length).  Waveforms and frame counts equal
    the exact two-graph path; an utterance whose prediction exceeds the bound is re-run exactly by synthesize_many."""
    from artspeech_b200 import engine
    model, g, _, _ = model_gpu
    gen = util.generator(0).to(DEV)
    gsrc = torch.Generator().manual_seed(31)
    tl = [40, 25, 33, 12]
    toks = [torch.randint(1, 178, (t,), generator=gsrc) for t in tl]
    mels = [(torch.randn(80, 110, generator=gsrc) * 0.5).clamp(-2, 2) for _ in tl]
    exact = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True)
    w_ref, f_ref = engine.synthesize_many(exact, toks, mels, None, max_batch=4, to_host=True)
    for bound in (4.0, 0.25):                                  # 0.25 frames per token: the longer utterances overflow
        syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, duration_bound=bound, pipeline_depth=2,
                                 frame_quantum=40 if bound == 4.0 else 8)
        for _ in range(2):
            w, f = engine.synthesize_many(syn, toks, mels, None, max_batch=4, to_host=True)
        assert f == f_ref, (bound, f, f_ref)
        for i in range(len(tl)):
            assert w[i].numel() == 300 * f[i]
            assert util.snr_db(w[i].float(), w_ref[i].float()) >= 50.0, (bound, i)
        if bound == 4.0:
            assert syn.stats["eager"] == 0 and syn.stats["captures"] <= 4
            # direct call: sync-free, lengths on the device, zeros beyond each utterance's length
            tok = torch.zeros(2, 40, dtype=torch.long); tok[0] = toks[0]; tok[1, :25] = toks[1]
            wav, lens, _ = syn.synthesize(tok.to(DEV), [40, 25], torch.stack(mels[:2]).to(DEV), [110, 110], None)
            syn.join()                                          # two batches in flight: the call ran on a slot stream
            assert syn.last["frames"] is None and lens.tolist() == f_ref[:2]
            assert float(wav[1, 300 * f_ref[1]:].abs().max()) == 0.0


def test_pipelined_step_soak(model_gpu):
    """300 back-to-back pipelined steps at the benchmark's shape (two graphs in flight, inputs changing every call) end
    without a sticky CUDA error and still give the first call's waveform for the first inputs.  Regression test for the
    CTA-pair implicit GEMM's ring depth (with a 6-stage ring this loop died within ~200 steps)."""
    import bench
    from artspeech_b200 import engine
    model, g, sd, dist_cpu = model_gpu
    gen = util.generator(0).to(DEV)
    syn = engine.Synthesizer(model, gen, device=DEV, use_cuda_graph=True, pipeline_depth=2)
    sets = [bench.make_inputs(0, 16, i) for i in range(4)]
    dev_sets = [(t.to(DEV), m.to(DEV), d) for t, _, m, _, d in sets]
    tl, ml = sets[0][1], sets[0][3]
    first = None
    for i in range(300):
        t, m, d = dev_sets[i % 4]
        wav, _, _ = syn.synthesize(t, tl, m, ml, d, predict_durations=True)
        if i == 0:
            syn.join()
            first = wav.clone()
        if i % 50 == 49:
            torch.cuda.synchronize()
    t, m, d = dev_sets[0]
    wav, _, _ = syn.synthesize(t, tl, m, ml, d, predict_durations=True)
    syn.join()
    torch.cuda.synchronize()
    assert torch.equal(wav, first)
