"""Same-run A/B of the whole vocoder (16 x 800 frames, graph replay) under different ASB_* environment settings, which
the C library reads at launch (= capture) time.  python tools/ab_vocoder.py "ASB_PAIR_2CTA=0" "ASB_PAIR_2CTA=1" ...
(the empty string "" = defaults).  Boxes differ by +-5 %: only numbers of one run compare."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import checkpoint

dev = torch.device("cuda:0")
gen = checkpoint.build_random_generator(0).to(dev).eval()
g0 = torch.Generator().manual_seed(0)
mel = torch.randn(16, 800, 80, generator=g0).clamp(-2, 2).to(dev).to(gen.compute_dtype)
lens = torch.full((16,), 800, dtype=torch.int32, device=dev)
if os.environ.get("RAGGED"):        # utterance lengths uniform in [50 %, 100 %] of the padded length, one full-length item
    lens = torch.randint(400, 801, (16,), generator=g0).to(torch.int32)
    lens[0] = 800
    print("ragged lengths, valid fraction %.3f" % (lens.float().mean().item() / 800))
    lens = lens.to(dev)
variants = sys.argv[1:] or [""]
graphs = []
for v in variants:
    sets = dict(kv.split("=") for kv in v.split(",") if kv)
    for k in [k for k in os.environ if k.startswith("ASB_PAIR") or k.startswith("ASB_ADAIN") or k.startswith("ASB_ATTN")]:
        del os.environ[k]
    os.environ.update(sets)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        gen.forward_channels_last(mel, lens)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = gen.forward_channels_last(mel, lens)
    graphs.append((v, g, out))
res = {v: [] for v in variants}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rnd in range(4):
    for v, g, _ in graphs:
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / 10)
flops = 623.7e6 * 16 * 800
for v in variants:
    ms = sorted(res[v])[len(res[v]) // 2]
    print(f"{v or 'defaults':40s} {ms:7.3f} ms   {flops / ms / 1e9:7.1f} TFLOP/s   (rounds: {' '.join(f'{x:.3f}' for x in res[v])})")
