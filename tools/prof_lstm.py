"""as_bilstm in a loop for ncu captures: python tools/prof_lstm.py H B T"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
xp = torch.randn(B, T, 8 * H, device=dev)
whh_t = (torch.randn(2, H, 4 * H, device=dev) / H ** 0.5).contiguous()
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.bilstm(xp, whh_t, H, lens, torch.float16)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.bilstm(xp, whh_t, H, lens, torch.float16)
e1.record(); torch.cuda.synchronize()
print(f"bilstm H={H} B={B} T={T}: {e0.elapsed_time(e1)/5*1e3:.1f} us  ({e0.elapsed_time(e1)/5*1e3/T:.2f} us/step)")
