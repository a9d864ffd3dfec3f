"""In-graph time of each phase of the acoustic model on the bench workload (CUDA-graph replay, CUDA events):
python tools/prof_phases.py"""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from artspeech_b200 import checkpoint, ops

dev = torch.device("cuda:0")
model = checkpoint.build_random_artsspeech(0).to(dev).eval()
model.distribution = {k: v.to(dev) for k, v in model.distribution.items()}
gen = checkpoint.build_random_generator(0).to(dev).eval()
tokens, tok_lens, mels, mel_lens, dur = bench.make_inputs(0)
tok_d, mel_d, dur_d = tokens.to(dev), mels.to(dev), dur.to(dev)
tl_d, ml_d = tok_lens.to(dev), mel_lens.to(dev)
meta = {"mel_lens": [int(v) for v in mel_lens], "Lmax": int(dur[0].sum())}
dt = model.compute_dtype


def timed(name, fn, reps=20):
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l0 = ops.launch_count
    with torch.cuda.graph(g):
        out = fn()
    n = ops.launch_count - l0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us   {n:4d} launches", flush=True)
    return out


with torch.no_grad():
    full = lambda: model([tok_d, tl_d, mel_d, ml_d], step="test", durations=dur_d, return_aux=True, host_meta=meta,
                         predict_durations=True)
    mel, aux = timed("acoustic model (all)", full)
    timed("durationPredictor", lambda: model.durationPredictor(tok_d, aux["ema_ext"], tl_d.to(torch.int32), ml_d,
                                                                host_mel_lengths=meta["mel_lens"]))
    timed("text_encoder", lambda: model.text_encoder(tok_d, tl_d))
    timed("arts_encoder", lambda: model.arts_encoder(tok_d, tl_d))
    timed("style_encoder", lambda: model.style_encoder(mel_d, ml_d, "second", model.distribution, host_lengths=meta["mel_lens"]))
    se = model.style_encoder
    mel_cl = ops.to_channels_last(mel_d, dt)
    B, M, T = mel_d.shape
    timed("  JDCNet", lambda: se.pitch_extractor.forward_cl(mel_cl.view(B, T, M, 1)))
    feat = torch.zeros(B, T, 88, dtype=dt, device=dev)
    timed("  EMA predictor", lambda: se.ema_extractor.forward_cl(feat[..., :82]))
    # predictor / decoder on the tensors of a full pass
    lens_t = tl_d.to(torch.int32)
    dur32 = dur_d.to(torch.int32).contiguous()
    style16 = aux["style"].to(dt).contiguous()
    Lmax = meta["Lmax"]
    a_reg, lens_l = ops.length_regulate(aux["A_en"], dur32, lens_t, 1, Lmax, out_dtype=model.artsPredictor.res_dtype)
    f0, n, ema, _ = timed("artsPredictor", lambda: model.artsPredictor.forward_cl(a_reg, style16, lens_l))
    D = model.decoder.dec_dim

    def dec():
        cat_res, cat16 = model.decoder.alloc_inputs(B, 2 * Lmax, dev)
        _, lens_m = ops.length_regulate(aux["T_en"], dur32, lens_t, 2, 2 * Lmax, out=cat_res[..., :D])
        if cat16 is not cat_res:
            ops.length_regulate(aux["T_en"], dur32, lens_t, 2, 2 * Lmax, out=cat16[..., :D])
        return model.decoder.forward_cl(cat_res, cat16, style16, f0, n, ema, lens_m)
    timed("length_regulate + decoder", dec)
    lens_m = aux["mel_lengths"]
    timed("vocoder", lambda: gen(mel, lens_m))
