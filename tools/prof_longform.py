"""BASELINE config 4 (long-form): 8 utterances x 60 s (900 tokens, 240-frame reference mel, sum(dur) = 2400 -> 4800 mel
frames each), full text -> waveform on one GPU through the CUDA-graph engine.  Prints one JSON line.
    python tools/prof_longform.py"""
import json, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import checkpoint, engine

B, TT, TR, SUM_DUR = 8, 900, 240, 2400
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(4321)
tokens = torch.randint(1, 178, (B, TT), generator=g)
mels = (torch.randn(B, 80, TR, generator=g) * 0.5).clamp(-2, 2)
base = SUM_DUR // TT
dur = torch.full((B, TT), base, dtype=torch.int64)
dur[:, : SUM_DUR - base * TT] += 1
tok_lens = torch.full((B,), TT, dtype=torch.int64)
mel_lens = torch.full((B,), TR, dtype=torch.int64)
syn = engine.Synthesizer(checkpoint.build_random_artsspeech(0), checkpoint.build_random_generator(0), device=dev,
                         use_cuda_graph=True, pipeline_depth=1)
tok_d, mel_d = tokens.to(dev), mels.to(dev)
for _ in range(3):
    wav, lens, mel = syn.synthesize(tok_d, tok_lens, mel_d, mel_lens, dur)
torch.cuda.synchronize()
steps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    wav, lens, mel = syn.synthesize(tok_d, tok_lens, mel_d, mel_lens, dur)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
audio_s = B * 2 * SUM_DUR * engine.HOP / engine.SAMPLE_RATE
assert wav.shape == (B, 2 * SUM_DUR * engine.HOP) and bool(torch.isfinite(wav).all())
print(json.dumps({"workload": "longform_8x60s_text2wave", "utterances": B, "tokens": TT, "mel_frames": 2 * SUM_DUR,
                  "audio_s_per_step": audio_s, "ms_per_step": round(ms, 3), "audio_s_per_s": round(audio_s / (ms / 1e3), 1),
                  "launches_per_step": syn.launches_per_call, "cuda_graph": True, "batches_in_flight": 1,
                  "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2)}))
