"""GPU dev check: MAS kernel vs the numpy oracle (bit-exact) + timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from artspeech_b200 import mas
from oracle import mas_oracle

torch.manual_seed(0)
rng = np.random.default_rng(0)
dev = "cuda"
ok_all = True
cases = [(1, 1, 1), (2, 5, 9), (4, 20, 50), (3, 33, 70), (8, 57, 130), (2, 200, 1000), (2, 257, 300), (1, 600, 1500), (1, 900, 2400)]
for (B, Tx, Ty) in cases:
    for kind in ("randn", "ties", "softmax"):
        if kind == "randn":
            v = rng.standard_normal((B, Tx, Ty)).astype(np.float32)
        elif kind == "ties":
            v = rng.integers(0, 3, (B, Tx, Ty)).astype(np.float32)
        else:
            z = torch.from_numpy(rng.standard_normal((B, Tx, Ty)).astype(np.float32) * 30)
            v = torch.softmax(z, dim=1).numpy()
        xl = rng.integers(1, Tx + 1, B); yl = np.maximum(rng.integers(max(Ty // 2, 1), Ty + 1, B), xl)
        yl = np.minimum(yl, Ty); xl = np.minimum(xl, yl)
        xl[0] = min(Tx, Ty); yl[0] = Ty
        if B > 1: xl[1] = 1
        for tie, name in ((0, "stay"), (1, "move")):
            ref = mas_oracle.maximum_path(v, xl, yl, name)
            vt = torch.from_numpy(v).to(dev)
            vt_copy = vt.clone()
            out = mas.maximum_path_lens(vt, torch.from_numpy(xl), torch.from_numpy(yl), tie)
            torch.cuda.synchronize()
            same = np.array_equal(out.cpu().numpy(), ref)
            unmod = torch.equal(vt, vt_copy)
            ok_all &= same and unmod
            if not (same and unmod):
                d = np.argwhere(out.cpu().numpy() != ref)
                print(f"MISMATCH B={B} Tx={Tx} Ty={Ty} {kind} {name}: ndiff={len(d)} first={d[:5].tolist()} unmod={unmod}")
print("MAS parity:", "OK" if ok_all else "FAIL")

# full-size config: 64 x 200 x 1000
B, Tx, Ty = 64, 200, 1000
v = torch.randn(B, Tx, Ty, device=dev)
xl = torch.randint(100, 201, (B,)); yl = torch.maximum(torch.randint(500, 1001, (B,)), xl)
xl[0] = 200; yl[0] = 1000
for _ in range(3):
    out = mas.maximum_path_lens(v, xl, yl, 0)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
xl_d = xl.to(dev, torch.int32); yl_d = yl.to(dev, torch.int32)
e0.record()
for _ in range(20):
    out = mas.maximum_path_lens(v, xl_d, yl_d, 0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
ref = mas_oracle.maximum_path(v.cpu().numpy(), xl.numpy(), yl.numpy(), "stay")
print(f"MAS 64x200x1000: {ms*1e3:.1f} us/call  ({2*B*Tx*Ty*4/ms/1e6:.1f} GB/s algorithmic)  exact={np.array_equal(out.cpu().numpy(), ref)}")
