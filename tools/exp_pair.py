"""Pair-kernel experiments: the shared-memory plan knobs of as_hifigan_resblock_pair (ASB_PAIR_* environment
overrides, read on every call) at the bench workload's stage sizes.  python tools/exp_pair.py [C ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops

KNOBS = ("ASB_PAIR_NX", "ASB_PAIR_NSTG", "ASB_PAIR_PIPE", "ASB_PAIR_SW", "ASB_PAIR_MODE", "ASB_PAIR_DBG", "ASB_PAIR_2CTA")


def run(C, L, k, dil, env, reps=5, B=16):
    for kn in KNOBS:
        os.environ.pop(kn, None)
    os.environ.update(env)
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    xa = (torch.randn(B, L, C, generator=g) * 0.5).to("cuda").to(dt)
    out = torch.empty_like(xa)
    w1 = torch.randn(k, C, C, generator=g) / (C * k) ** 0.5
    w2 = torch.randn(k, C, C, generator=g) / (C * k) ** 0.5
    b1, b2 = torch.randn(C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    p1 = ops.pack_conv(w1, b1, ops.taps_1d(k, dil), dt, "cuda")
    p2 = ops.pack_conv(w2, b2, ops.taps_1d(k, 1), dt, "cuda")
    f = lambda: ops.resblock_pair(xa, p1, p2, k, dil, slope=0.1, out_act=ops.ACT_LRELU, out_slope=0.1, out=out)
    try:
        for _ in range(2):
            f()
        torch.cuda.synchronize()
    except Exception as e:
        return None, str(e)[:60]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 2.0 * 2 * B * L * C * C * k / us / 1e6


def main():
    only = [int(v) for v in sys.argv[1:]]
    if os.environ.get("EXP") == "2cta":
        names = ["2cta", "1cta", "2cta", "1cta"]
        variants = [{"ASB_PAIR_2CTA": "1"}, {"ASB_PAIR_2CTA": "0"}, {"ASB_PAIR_2CTA": "1"}, {"ASB_PAIR_2CTA": "0"}]
    elif os.environ.get("EXP") == "ab":
        names = ["fast", "nofast", "fast", "nofast"]
        variants = [{}, {"ASB_PAIR_DBG": "64"}, {}, {"ASB_PAIR_DBG": "64"}]
    elif os.environ.get("EXP") == "dbg":
        # timing decomposition (results are wrong): what is left when a part of the tile pipeline is switched off
        names = ["all", "noMMA", "noStore", "noReload", "noEpi1", "noSeed", "noMMA+noLdSt", "chainOnly"]
        variants = [{"ASB_PAIR_DBG": str(v)} for v in (0, 1, 2, 4, 8, 16, 7, 31)]
    else:
        names = ["default", "NX2", "NX3", "NX4", "PIPE0", "NSTG1", "NX4+NSTG1"]
        variants = [{}, {"ASB_PAIR_NX": "2"}, {"ASB_PAIR_NX": "3"}, {"ASB_PAIR_NX": "4"}, {"ASB_PAIR_PIPE": "0"},
                    {"ASB_PAIR_NSTG": "1"}, {"ASB_PAIR_NX": "4", "ASB_PAIR_NSTG": "1"}]
    for C, L in ((32, 240000), (64, 120000), (128, 40000)):
        if only and C not in only:
            continue
        for k, dil in ((3, 1), (3, 5), (7, 3), (11, 1), (11, 5)):
            row = []
            for env in variants:
                us, tf = run(C, L, k, dil, env)
                row.append("   fail" if us is None else f"{us:7.1f}")
            print(f"C={C:3d} k={k:2d} d={dil}: " + " ".join(row) + "   us  [" + " ".join(names) + "]", flush=True)


if __name__ == "__main__":
    main()
