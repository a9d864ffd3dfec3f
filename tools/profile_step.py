"""One eager (no CUDA graph) step of the bench workload, for `ncu` launch lists.
Usage: ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> ... python tools/profile_step.py [--vocoder-only]
Prints the number of our launches per step so -s/-c can be set."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from artspeech_b200 import checkpoint, engine, ops

dev = torch.device("cuda:0")
syn = engine.Synthesizer(checkpoint.build_random_artsspeech(0), checkpoint.build_random_generator(0), device=dev,
                         use_cuda_graph=False)
tokens, tok_lens, mels, mel_lens, dur = bench.make_inputs(0)
tok_d, mel_d = tokens.to(dev), mels.to(dev)
steps = int(os.environ.get("STEPS", "2"))
for i in range(steps):
    l0 = ops.launch_count
    torch.cuda.synchronize()
    if "--vocoder-only" in sys.argv and i > 0:
        syn.generator(mel_static, lens_static)
    else:
        wav, lens_static, mel_static = syn.synthesize(tok_d, tok_lens, mel_d, mel_lens, dur)
        mel_static = mel_static.clone()
    torch.cuda.synchronize()
    print(f"step {i}: {ops.launch_count - l0} launches", flush=True)
