"""Single conv shape in a loop, for `ncu --set full` captures: python tools/prof_conv.py C k rows [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artspeech_b200 import ops
C, k, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
use_res = len(sys.argv) > 4 and sys.argv[4] == "res"
dev = "cuda"
dt = torch.bfloat16
B = 16
x = torch.randn(B, rows // B, C, device=dev).to(dt)
w = torch.randn(k, C, C) / (C * k) ** 0.5
pw = ops.pack_conv(w, torch.zeros(C), ops.taps_1d(k, 1), dt, dev)
out = torch.empty_like(x)
for _ in range(6):
    ops.conv(x, pw, res1=x if use_res else None, act_out=out, act=ops.ACT_LRELU, slope=0.1)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv(x, pw, res1=x if use_res else None, act_out=out, act=ops.ACT_LRELU, slope=0.1)
e1.record(); torch.cuda.synchronize()
print(f"C={C} k={k} rows={rows} res={use_res}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
