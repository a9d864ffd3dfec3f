"""GPU dev check: as_conv_igemm vs torch conv on the same 16-bit-rounded operands."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as Fn
from artspeech_b200 import ops

dev = "cuda"
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

def check1d(B, T, Cin, Cout, k, dil, dt, res=False, lens=None, act=ops.ACT_NONE, slope=0.1, raw_dt=torch.float32, tag=""):
    x = (torch.randn(B, T, Cin, device=dev) * 0.5).to(dt)
    w = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
    bias = torch.randn(Cout, device=dev)
    wq = w.to(dt).float()
    pw = ops.pack_conv(w.permute(2, 0, 1).contiguous(), bias, ops.taps_1d(k, dil), dt, dev)
    r1 = torch.randn(B, T, Cout, device=dev).to(dt) if res else None
    lens_t = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=dev)
    raw, actt = ops.conv(x, pw, res1=r1, scale=0.5, raw=raw_dt, act_out=dt, act=act, slope=slope, lens=lens_t)
    torch.cuda.synchronize()
    pad = (k * dil - dil) // 2
    ref = Fn.conv1d(x.float().transpose(1, 2), wq, bias, padding=pad, dilation=dil).transpose(1, 2)
    if res: ref = ref + r1.float()
    ref = ref * 0.5
    if lens is not None:
        m = (torch.arange(T, device=dev)[None, :] < lens_t[:, None]).float()[..., None]
        ref = ref * m
    err = (raw.float() - ref).abs().max().item()
    ra = ref
    if act == ops.ACT_LRELU: ra = Fn.leaky_relu(ref, slope)
    elif act == ops.ACT_RELU: ra = torch.relu(ref)
    elif act == ops.ACT_TANH: ra = torch.tanh(ref)
    erra = (actt.float() - ra).abs().max().item()
    scale = ref.abs().max().item()
    status = "OK " if err < 2e-2 * max(scale, 1) and erra < 3e-2 * max(scale, 1) else "BAD"
    print(f"{status} conv1d{tag} B{B} T{T} {Cin}->{Cout} k{k} d{dil} {str(dt)[6:]} res={res} lens={lens is not None}: raw_err={err:.3e} act_err={erra:.3e} (max|ref|={scale:.2f})", flush=True)
    return status == "OK "

def check2d(B, T, F, Cin, Cout, dt):
    x = (torch.randn(B, T, F, Cin, device=dev) * 0.5).to(dt)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5   # [co, ci, kt, kf]
    wq = w.to(dt).float()
    wt = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    pw = ops.pack_conv(wt, None, ops.taps_2d(3, 3, 1, 1), dt, dev)
    raw, _ = ops.conv(x, pw, raw=torch.float32)
    torch.cuda.synchronize()
    ref = Fn.conv2d(x.float().permute(0, 3, 1, 2), wq, None, padding=1).permute(0, 2, 3, 1)
    err = (raw - ref).abs().max().item()
    status = "OK " if err < 2e-2 else "BAD"
    print(f"{status} conv2d B{B} T{T} F{F} {Cin}->{Cout} {str(dt)[6:]}: err={err:.3e}", flush=True)

bf, hf = torch.bfloat16, torch.float16
print("device:", torch.cuda.get_device_name(0), flush=True)
# smallest sanity first
check1d(1, 128, 64, 64, 1, 1, bf, tag="[gemm]")
check1d(1, 128, 64, 64, 3, 1, bf)
check1d(2, 300, 128, 128, 3, 1, bf)
check1d(2, 300, 128, 128, 7, 3, bf, res=True, act=ops.ACT_LRELU)
check1d(2, 300, 256, 256, 11, 5, bf, res=True, act=ops.ACT_LRELU, lens=[300, 177])
check1d(2, 1000, 32, 32, 3, 1, bf, res=True, act=ops.ACT_LRELU)
check1d(2, 1000, 32, 32, 11, 5, bf)
check1d(2, 1000, 64, 64, 7, 3, bf, raw_dt=bf)
check1d(1, 200, 80, 512, 7, 1, bf)
check1d(1, 900, 32, 1, 7, 1, bf, act=ops.ACT_TANH)
check1d(2, 150, 512, 1536, 1, 1, hf)
check1d(2, 150, 512, 1024, 9, 1, hf, act=ops.ACT_RELU, lens=[150, 77])
check1d(2, 400, 1216, 1024, 3, 1, hf)
check1d(2, 400, 640, 1024, 3, 1, hf, res=True)
check1d(1, 800, 512, 80, 1, 1, hf)
check1d(1, 77, 512, 2560, 3, 1, bf, tag="[ups0-shape]")
check1d(3, 50, 82 + 6, 256, 1, 1, hf, tag="[Cin=88]")
check2d(1, 64, 80, 64, 64, hf)
check2d(2, 100, 40, 64, 128, hf)
check2d(1, 60, 5, 256, 512, hf)
check2d(1, 240, 10, 192, 256, hf)

# timing of vocoder-like shapes (B=16, T=800 mel frames)
def bench(B, T, C, k, dil, dt=bf, iters=20):
    x = torch.randn(B, T, C, device=dev).to(dt)
    w = torch.randn(k, C, C) / (C * k) ** 0.5
    pw = ops.pack_conv(w, torch.zeros(C), ops.taps_1d(k, dil), dt, dev)
    out = torch.empty(B, T, C, device=dev, dtype=dt)
    for _ in range(3): ops.conv(x, pw, res1=x, act_out=out, act=ops.ACT_LRELU, slope=0.1)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.conv(x, pw, res1=x, act_out=out, act=ops.ACT_LRELU, slope=0.1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * T * C * C * k
    by = 3.0 * B * T * C * 2
    print(f"time conv C={C} k={k} d={dil} rows={B*T}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  {by/ms/1e6:.0f} GB/s(alg)", flush=True)

for (C, L) in ((256, 10), (128, 50), (64, 150), (32, 300)):
    for k in (3, 7, 11):
        bench(16, 800 * L, C, k, 1)
