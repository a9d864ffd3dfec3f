"""Time one conv shape with each epilogue kind (16-bit TMA epilogue; fp32 raw; fp32 residual + fp32 raw + f16 act) to
see what the direct (non-TMA) epilogue costs.  python tools/prof_epilogue_kinds.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops
dev = "cuda"
B = 16
torch.manual_seed(0)
for (cin, cout, k, T) in [(512, 512, 3, 800), (1024, 1024, 3, 800), (1216, 1024, 3, 800), (512, 512, 1, 150), (512, 1024, 9, 150)]:
    x = torch.randn(B, T, cin, device=dev).to(torch.float16)
    p = ops.pack_conv(torch.randn(k, cout, cin) / (cin * k) ** 0.5, torch.zeros(cout), ops.taps_1d(k, 1), torch.float16, dev)
    r = torch.randn(B, T, cout, device=dev)
    o = torch.empty(B, T, cout, device=dev)
    o16 = torch.empty(B, T, cout, device=dev, dtype=torch.float16)
    r16 = r.to(torch.float16)
    l = torch.full((B,), T, dtype=torch.int32, device=dev)
    kinds = {
        "act16 (TMA)": lambda: ops.conv(x, p, act_out=o16, act=ops.ACT_LRELU, slope=0.2, lens=l),
        "res16+raw16 (TMA)": lambda: ops.conv(x, p, res1=r16, raw=o16, lens=l),
        "raw32": lambda: ops.conv(x, p, raw=o, lens=l),
        "res32+raw32": lambda: ops.conv(x, p, res1=r, raw=o, lens=l),
        "res32+raw32+act16": lambda: ops.conv(x, p, res1=r, raw=o, act_out=o16, act=ops.ACT_LRELU, slope=0.2, lens=l),
    }
    flops = 2.0 * B * T * cin * cout * k
    line = []
    for name, f in kinds.items():
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        line.append(f"{name} {us:6.1f} us ({flops / us / 1e6:5.0f} TF/s)")
    print(f"{cin}->{cout} k{k} T={T}: " + " | ".join(line), flush=True)
