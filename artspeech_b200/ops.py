"""Host-side wrappers of the C-ABI primitives (``include/artspeech_b200.h``).

Everything here is plumbing: pointer extraction from torch CUDA tensors, output allocation through
torch's caching allocator, the current stream.  The arithmetic happens in the .so.  Activations are
channels-last (``[B, T, C]`` / ``[B, T, F, C]``); a channel slice of a wider buffer is passed as a
strided torch view (``stride(-1) == 1``, row stride = the buffer's channel count), which is how
concatenations are fused away.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ACT_ABS, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SWISH, ACT_TANH, AS_BF16, AS_F16,
                   AS_F32)

_DT = {torch.float16: AS_F16, torch.bfloat16: AS_BF16, torch.float32: AS_F32}

# number of kernels of ours launched since the last reset (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise _lib.AsError(f"unsupported dtype {dt}") from None


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.AsError(f"{name}: artspeech_b200 ops need CUDA tensors (there is no CPU fallback)")


def _rows_ld(t: torch.Tensor, name: str) -> int:
    """Row stride (elements) of a channels-last tensor/view whose outer dims are dense."""
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise _lib.AsError(f"{name}: channel stride must be 1")
    ld = t.stride(-2) if t.dim() >= 2 else t.shape[-1]
    exp = ld
    for d in range(t.dim() - 2, 0, -1):
        exp *= t.shape[d]
        if t.shape[d - 1] != 1 and t.stride(d - 1) != exp:
            raise _lib.AsError(f"{name}: outer dims must be dense (shape {tuple(t.shape)}, strides {t.stride()})")
    return ld


# --------------------------------------------------------------------------------------------
# implicit-GEMM convolution
# --------------------------------------------------------------------------------------------
@dataclass
class PackedConv:
    """Weights of one dense contraction in the kernel's layout ``[ntaps][CoutP][CinP]`` (K-major)."""
    w: torch.Tensor                 # 16-bit, device
    bias: Optional[torch.Tensor]    # fp32 [Cout], device (or None)
    ntaps: int
    Cin: int
    Cout: int
    CinP: int
    CoutP: int
    taps: Sequence[tuple]           # (dt, df) per tap
    _dt: object = field(default=None, repr=False)
    _df: object = field(default=None, repr=False)

    def __post_init__(self):
        n = len(self.taps)
        self._dt = (C.c_int32 * n)(*[int(t[0]) for t in self.taps])
        self._df = (C.c_int32 * n)(*[int(t[1]) for t in self.taps])


def conv_tile_n(cout: int) -> int:
    """Same rule as ``as_conv_tile_n`` (kept in Python so packing works without a GPU)."""
    if cout <= 16:
        return 16
    if cout <= 32:
        return 32
    if cout <= 64:
        return 64
    if cout <= 128:
        return 128
    if cout % 256 == 0:
        return 256
    if cout % 128 == 0:
        return 128
    p256 = (cout + 255) // 256 * 256
    p128 = (cout + 127) // 128 * 128
    return 128 if p128 < p256 else 256


def pack_conv(w_taps: torch.Tensor, bias: Optional[torch.Tensor], taps: Sequence[tuple],
              dtype: torch.dtype, device) -> PackedConv:
    """``w_taps`` fp32 ``[ntaps, Cout, Cin]`` -> zero-padded 16-bit ``[ntaps, CoutP, CinP]``."""
    ntaps, cout, cin = w_taps.shape
    assert ntaps == len(taps)
    bk = 64 if cin >= 64 else 32
    cinp = (cin + bk - 1) // bk * bk
    bn = conv_tile_n(cout)
    coutp = (cout + bn - 1) // bn * bn
    wp = torch.zeros(ntaps, coutp, cinp, dtype=torch.float32)
    wp[:, :cout, :cin] = w_taps.detach().float().cpu()
    wp = wp.to(dtype).to(device).contiguous()
    b = None if bias is None else bias.detach().float().to(device).contiguous()
    return PackedConv(wp, b, ntaps, cin, cout, cinp, coutp, list(taps))


def _out_arg(spec, shape, device):
    """spec: None | torch.dtype (allocate) | tensor/view (write into it)."""
    if spec is None:
        return None
    if isinstance(spec, torch.dtype):
        return torch.empty(shape, dtype=spec, device=device)
    return spec


def conv(x: torch.Tensor, pw: PackedConv, *, out_shape=None, use_bias: bool = True,
         res1: Optional[torch.Tensor] = None, res2: Optional[torch.Tensor] = None,
         scale: float = 1.0, raw=None, act_out=None, act: int = ACT_NONE, slope: float = 0.0,
         lens: Optional[torch.Tensor] = None):
    """Run ``as_conv_igemm``.

    ``x``: ``[B, T, C]`` or ``[B, T, F, C]`` 16-bit channels-last (may be a channel-slice view).
    ``out_shape``: ``(To, Fo)`` when the output grid differs from the input's ('valid' convs).
    ``raw`` / ``act_out``: ``None``, a dtype (allocate) or a tensor/view to write into.
    Returns ``(raw_tensor_or_None, act_tensor_or_None)``.
    """
    _require_cuda(x, "conv")
    lib = _lib.load()
    if x.dim() == 3:
        B, T, Cin = x.shape
        F = 1
    else:
        B, T, F, Cin = x.shape
    if Cin != pw.Cin:
        raise _lib.AsError(f"conv: input has {Cin} channels, weights expect {pw.Cin}")
    To, Fo = (T, F) if out_shape is None else out_shape
    oshape = (B, To, pw.Cout) if x.dim() == 3 else (B, To, Fo, pw.Cout)
    raw_t = _out_arg(raw, oshape, x.device)
    act_t = _out_arg(act_out, oshape, x.device)
    p = _lib.ConvParams()
    p.x = x.data_ptr(); p.x_dtype = dtype_code(x.dtype)
    p.B, p.T, p.F, p.Cin = B, T, F, Cin
    p.x_ld = _rows_ld(x, "conv.x")
    if pw.w.dtype != x.dtype:
        raise _lib.AsError(f"conv: x is {x.dtype} but weights were packed as {pw.w.dtype}")
    p.w = pw.w.data_ptr()
    p.ntaps, p.CinP, p.CoutP, p.Cout = pw.ntaps, pw.CinP, pw.CoutP, pw.Cout
    p.tap_dt = C.cast(pw._dt, _lib.c_i32_p); p.tap_df = C.cast(pw._df, _lib.c_i32_p)
    p.To, p.Fo = To, Fo
    p.bias = pw.bias.data_ptr() if (use_bias and pw.bias is not None) else None
    for name, r in (("res1", res1), ("res2", res2)):
        if r is not None:
            setattr(p, name, r.data_ptr())
            setattr(p, name + "_dtype", dtype_code(r.dtype))
            setattr(p, name + "_ld", _rows_ld(r, "conv." + name))
    p.out_scale = float(scale)
    if raw_t is not None:
        p.y_raw = raw_t.data_ptr(); p.y_raw_dtype = dtype_code(raw_t.dtype)
        p.y_raw_ld = _rows_ld(raw_t, "conv.raw")
    if act_t is not None:
        p.y_act = act_t.data_ptr(); p.y_act_dtype = dtype_code(act_t.dtype)
        p.y_act_ld = _rows_ld(act_t, "conv.act_out")
    p.act = int(act); p.slope = float(slope)
    if lens is not None:
        if lens.dtype != torch.int32:
            raise _lib.AsError("conv: lens must be int32")
        p.lens = lens.data_ptr()
    with torch.cuda.device(x.device):
        rc = lib.as_conv_igemm(C.byref(p), _stream(x))
    _lib.check(rc, "as_conv_igemm")
    _count()
    return raw_t, act_t


# tap tables -------------------------------------------------------------------------------------
def taps_1d(k: int, dilation: int = 1, padding: Optional[int] = None):
    """Conv1d taps: input index = t + j*dilation - padding ('same' padding by default)."""
    if padding is None:
        padding = (k * dilation - dilation) // 2
    return [(j * dilation - padding, 0) for j in range(k)]


def taps_2d(kt: int, kf: int, pad_t: int, pad_f: int):
    """Conv2d taps over (T, F); order matches ``w.permute`` in ``conv2d_weight_taps``."""
    return [(jt - pad_t, jf - pad_f) for jt in range(kt) for jf in range(kf)]
