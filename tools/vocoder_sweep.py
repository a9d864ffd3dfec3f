"""BASELINE config 3: vocoder-only sweep, mel 80 x {200, 800, 3200} frames at batch 1 .. 64 (bf16 operands, fp32
accumulate), CUDA-graph replay, CUDA events.  FLOPs = 623.7e6 per mel frame.  Prints one JSON line
per point and a summary table on stderr.
    python tools/vocoder_sweep.py"""
import json, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import checkpoint

FLOP_PER_FRAME = 623.7e6
dev = torch.device("cuda:0")
gen = checkpoint.build_random_generator(0).to(dev).eval()
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
rows = []
for T in (200, 800, 3200):
    for B in (1, 2, 4, 8, 16, 32, 64):
        torch.manual_seed(0)
        mel = torch.randn(B, 80, T, device=dev).clamp(-2, 2)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                gen(mel)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            wav = gen(mel)
        g.replay(); torch.cuda.synchronize()
        reps = 5 if B * T >= 12800 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = FLOP_PER_FRAME * B * T / (ms * 1e-3) / 1e12
        row = {"workload": "vocoder_sweep", "B": B, "T": T, "frames": B * T, "ms": round(ms, 3), "tflops": round(tf, 1),
               "audio_s_per_s": round(B * T * 300 / 24000 / (ms * 1e-3), 1)}
        if peak:
            row["frac_of_sustained_bf16_peak"] = round(tf / peak["bf16_tflops_sustained"], 3)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del g, wav, mel
        torch.cuda.empty_cache()
print("T \\ B   " + "".join(f"{b:>8d}" for b in (1, 2, 4, 8, 16, 32, 64)), file=sys.stderr)
for T in (200, 800, 3200):
    print(f"{T:6d}   " + "".join(f"{r['tflops']:8.0f}" for r in rows if r["T"] == T), file=sys.stderr)
