"""Achieved HBM bandwidth of the fused InstanceNorm + AdaIN + LeakyReLU kernel (as_adain_norm_apply) at the
decoder's shapes: python tools/prof_adain.py   -> us, algorithmic GB/s = (read x + write out) / time"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops

def run(B, T, C, xdt, up=False, reps=5, nset=6):
    """Rotates over `nset` distinct input / output sets (> 126 MB L2 in total) and times `reps` rounds of
    back-to-back launches with one pair of CUDA events: no launch gaps, no L2 reuse, no dirty-line flush."""
    xs = [(torch.randn(B, T, C, device="cuda") * 0.7 + 3.0).to(xdt) for _ in range(nset)]
    gb = torch.randn(B, 2 * C, device="cuda") * 0.3
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    up_w = torch.randn(C, 3, device="cuda") if up else None
    up_b = torch.randn(C, device="cuda") if up else None
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for x in xs:
            ops.adain_norm(x, gb, 0.2, lens, torch.float16, up_w, up_b)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        outs = [ops.adain_norm(x, gb, 0.2, lens, torch.float16, up_w, up_b) for x in xs]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (reps * nset) * 1e3
    byts = B * T * C * (xs[0].element_size() + 2 * (2 if up else 1))
    print(f"B={B} T={T} C={C} x={str(xdt)[6:]} up={up}: {us:7.1f} us  {byts / us / 1e3:7.1f} GB/s ({byts / 1e6:.0f} MB per launch, "
          f"{nset} rotating sets)", flush=True)
    return us, byts


if __name__ == "__main__":
    for C in (512, 1024, 1216):
        run(16, 800, C, torch.float32)
    run(512, 800, 32, torch.float32)      # same bytes as 16x800x1024 but every CTA's slab is contiguous (DRAM-page experiment)
    run(16, 400, 512, torch.float32, up=True)
    run(16, 1600, 512, torch.float32)     # long path (stats + apply)
