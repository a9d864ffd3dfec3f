"""Dev check of as_hifigan_resblock_pair on a GPU box: python tools/dev_check_pair.py [C k dil L B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops
from tests import sim_backend as sim

def run(C, k, dil, L, B, dt=torch.bfloat16):
    torch.manual_seed(0)
    x_raw = torch.randn(B, L, C)
    xa = torch.where(x_raw > 0, x_raw, 0.1 * x_raw).to(dt)
    w1 = torch.randn(k, C, C) / (C * k) ** 0.5
    w2 = torch.randn(k, C, C) / (C * k) ** 0.5
    b1, b2 = torch.randn(C) * 0.1, torch.randn(C) * 0.1
    pc = (ops.pack_conv(w1, b1, ops.taps_1d(k, dil), dt, "cpu"), ops.pack_conv(w2, b2, ops.taps_1d(k, 1), dt, "cpu"))
    pg = (ops.pack_conv(w1, b1, ops.taps_1d(k, dil), dt, "cuda"), ops.pack_conv(w2, b2, ops.taps_1d(k, 1), dt, "cuda"))
    ref = sim.resblock_pair(xa, *pc, k, dil, slope=0.1, out_act=ops.ACT_LRELU, out_slope=0.1).float()
    out = ops.resblock_pair(xa.cuda(), *pg, k, dil, slope=0.1, out_act=ops.ACT_LRELU, out_slope=0.1)
    torch.cuda.synchronize()
    out = out.float().cpu()
    err = (out - ref).abs()
    bad = (err > 0.05).nonzero()
    print(f"C={C} k={k} d={dil} L={L} B={B}: max err {err.max():.4f} mean {err.mean():.5f} ref absmax {ref.abs().max():.2f} bad {len(bad)}", flush=True)
    if len(bad):
        rows = sorted(set(int(r[1]) for r in bad))
        print("   bad rows (first 20):", rows[:20], "... last", rows[-5:], " cols:", sorted(set(int(r[2]) for r in bad))[:16])

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(*[int(v) for v in sys.argv[1:6]])
    else:
        for cfg in [(64, 3, 1, 700, 1), (64, 7, 3, 1000, 2), (32, 3, 1, 700, 1), (32, 11, 5, 2000, 2), (64, 11, 5, 1500, 1),
                    (128, 3, 1, 700, 1), (128, 7, 3, 900, 2), (128, 11, 5, 700, 1)]:
            run(*cfg)
