"""One launch of each tensor-core attention kernel at the bench workload's shapes, for ncu captures:
python tools/prof_attention.py [relpos|conformer]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "relpos"
dev = "cuda"
torch.manual_seed(0)
if which == "relpos":
    B, T, H, D = 16, 160, 4, 128
    qkv = torch.randn(B, T, 3 * H * D, device=dev)
    ek, ev = torch.randn(9, D, device=dev) * D ** -0.5, torch.randn(9, D, device=dev) * D ** -0.5
    lens = torch.full((B,), 150, dtype=torch.int32, device=dev)
    f = lambda: ops.relpos_attention(qkv, ek, ev, 4, H, lens, torch.float16)
else:
    B, T, H, D = 16, 240, 4, 64
    qkv = torch.randn(B, T, 3 * H * D, device=dev)
    pos = torch.randn(T, H * D, device=dev)
    u, v = torch.randn(H, D, device=dev) * 0.1, torch.randn(H, D, device=dev) * 0.1
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    f = lambda: ops.conformer_attention(qkv[..., :256], qkv[..., 256:512], qkv[..., 512:], pos, u, v, H, lens, torch.float16)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    f()
e1.record(); torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (warm, back to back)")
