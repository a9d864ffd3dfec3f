"""Time as_hifigan_resblock_pair at the bench workload's stage sizes (16 utterances x 800 mel frames):
python tools/prof_pair.py [reps [C [k [dil]]]]   -> per (C, k, dil): us, TFLOP/s, algorithmic GB/s"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    only = [int(v) for v in sys.argv[2:3]]
    only_k = int(sys.argv[3]) if len(sys.argv) > 3 else None
    only_d = int(sys.argv[4]) if len(sys.argv) > 4 else None
    dt = torch.bfloat16
    B = 16
    tot = 0.0
    for C, L in ((128, 40000), (64, 120000), (32, 240000)):
        if only and C not in only:
            continue
        xa = (torch.randn(B, L, C, device="cuda") * 0.5).to(dt)
        out = torch.empty_like(xa)
        for k in (3, 7, 11):
            for dil in (1, 3, 5):
                if (only_k is not None and k != only_k) or (only_d is not None and dil != only_d):
                    continue
                w1 = torch.randn(k, C, C) / (C * k) ** 0.5
                w2 = torch.randn(k, C, C) / (C * k) ** 0.5
                b1, b2 = torch.randn(C) * 0.1, torch.randn(C) * 0.1
                p1 = ops.pack_conv(w1, b1, ops.taps_1d(k, dil), dt, "cuda")
                p2 = ops.pack_conv(w2, b2, ops.taps_1d(k, 1), dt, "cuda")
                f = lambda: ops.resblock_pair(xa, p1, p2, k, dil, slope=0.1, out_act=ops.ACT_LRELU, out_slope=0.1, out=out)
                for _ in range(2):
                    f()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    f()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / reps * 1e3
                flops = 2.0 * 2 * B * L * C * C * k
                byts = 2.0 * B * L * C * 2
                tot += us
                print(f"C={C:3d} k={k:2d} d={dil}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s (read+write once)", flush=True)
    print(f"sum over listed pairs: {tot / 1e3:.3f} ms")

if __name__ == "__main__":
    main()
