"""Time as_conv_igemm on the acoustic model's representative GEMM shapes (fp16, fp32 outputs):
python tools/prof_gemm_shapes.py   (compare with ASB_NO_2CTA=1 / ASB_BN128_MAX_TILES=0)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops

def run(name, B, T, Cin, Cout, k, out_dt=torch.float32, dt=torch.float16, reps=20):
    x = torch.randn(B, T, Cin, device="cuda").to(dt)
    w = torch.randn(k, Cout, Cin) / (Cin * k) ** 0.5
    pw = ops.pack_conv(w, torch.zeros(Cout), ops.taps_1d(k, 1), dt, "cuda")
    out = torch.empty(B, T, Cout, device="cuda", dtype=out_dt)
    f = lambda: ops.conv(x, pw, raw=out)
    for _ in range(3):
        f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:34s} M={B*T:6d} K={Cin*k:5d} N={Cout:5d}: {us:7.1f} us  {2.0*B*T*Cin*k*Cout/us/1e6:7.1f} TFLOP/s", flush=True)

run("decoder decode[0] conv1 (1216->1024 k3)", 16, 800, 1216, 1024, 3)
run("decoder decode conv2 (1024->1024 k3)", 16, 800, 1024, 1024, 3)
run("decoder encode conv1 (640->1024 k3)", 16, 800, 640, 1024, 3)
run("decoder 512->512 k3", 16, 800, 512, 512, 3)
run("predictor 512->512 k3 @L=400", 16, 400, 512, 512, 3)
run("text enc FFN 512->1024 k9", 16, 150, 512, 1024, 9)
run("text enc 1024->512 k1", 16, 150, 1024, 512, 1)
run("conformer FFN 256->1024", 16, 240, 256, 1024, 1)
run("vocoder stage0 256->256 k11 (bf16)", 16, 8000, 256, 256, 11, out_dt=torch.bfloat16, dt=torch.bfloat16)
run("vocoder ups0 512->2560 k3 (bf16)", 16, 800, 512, 2560, 3, out_dt=torch.bfloat16, dt=torch.bfloat16)
