"""Hand-off timeline of as_hifigan_resblock_pair (ASB_PAIR_DBG=32): clock64() of CTA 0 at every barrier hand-off of
its first tiles.  python tools/trace_pair.py C k dil [extra dbg bits]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import _lib, ops

NAMES = ["prod_buf", "S1_beg", "S1_end", "S2_beg", "S2_end", "g1_has", "seeded", "D1_rdy", "g1_sync", "epi1_done",
         "g1_sig", "g2_sees", "g2_done", "x_landed", "i2_entry", "i2_tfull", "i1_entry", "g2_ld", "g2_sts", "g2_tma"]


def main():
    Cc, k, dil = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    extra = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    L = {32: 240000, 64: 120000, 128: 40000}[Cc]
    dt = torch.bfloat16
    xa = (torch.randn(16, L, Cc, device="cuda") * 0.5).to(dt)
    out = torch.empty_like(xa)
    p1 = ops.pack_conv(torch.randn(k, Cc, Cc) / (Cc * k) ** 0.5, torch.randn(Cc) * 0.1, ops.taps_1d(k, dil), dt, "cuda")
    p2 = ops.pack_conv(torch.randn(k, Cc, Cc) / (Cc * k) ** 0.5, torch.randn(Cc) * 0.1, ops.taps_1d(k, 1), dt, "cuda")
    os.environ["ASB_PAIR_DBG"] = str(32 | extra)
    for _ in range(2):
        ops.resblock_pair(xa, p1, p2, k, dil, slope=0.1, out_act=ops.ACT_LRELU, out_slope=0.1, out=out)
    torch.cuda.synchronize()
    lib = _lib.load()
    buf = (C.c_ulonglong * (20 * 32))()
    lib.as_debug_pair_trace.argtypes = [C.c_void_p, C.c_int]
    assert lib.as_debug_pair_trace(buf, 20 * 32) == 0
    NE = len(NAMES)
    t = [[buf[e * 32 + i] for i in range(32)] for e in range(NE)]
    t0 = t[1][6]
    print(f"C={Cc} k={k} d={dil} dbg={32 | extra}: clocks relative to S1_beg of tile 6")
    print("tile " + " ".join(f"{n:>9s}" for n in NAMES))
    for i in range(6, 16):
        print(f"{i:4d} " + " ".join(f"{int(t[e][i]) - int(t0):9d}" for e in range(NE)))
    per = (t[4][25] - t[4][9]) / 16.0
    print(f"steady-state tile period: {per:.0f} clk")


if __name__ == "__main__":
    main()
