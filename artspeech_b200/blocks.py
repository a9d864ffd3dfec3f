"""Building blocks of ``models.py`` (parameter containers + kernel-level runners).

Containers mirror the reference's attribute names so checkpoints load unchanged
(models.py:19-270); the ``run_*`` functions are the channels-last dataflow on the C-ABI kernels.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm, weight_norm

from . import nn_util, ops

INV_SQRT2 = 1.0 / math.sqrt(2.0)
LRELU = 0.2


# --------------------------------------------------------------------------------------------
# AdaIN residual block, 1-D (models.py:158-240)
# --------------------------------------------------------------------------------------------
class AdaIN1d(nn.Module):
    def __init__(self, style_dim: int, num_features: int):
        super().__init__()
        self.num_features = num_features
        self.fc = nn.Linear(style_dim, num_features * 2)


class AdainResBlk1d(nn.Module):
    def __init__(self, dim_in, dim_out, style_dim=64, upsample="none", dropout_p=0.0):
        super().__init__()
        self.dim_in, self.dim_out, self.style_dim = dim_in, dim_out, style_dim
        self.upsample_type = upsample
        self.is_up = upsample != "none"
        self.learned_sc = dim_in != dim_out
        self.conv1 = weight_norm(nn.Conv1d(dim_in, dim_out, 3, 1, 1))
        self.conv2 = weight_norm(nn.Conv1d(dim_out, dim_out, 3, 1, 1))
        self.norm1 = AdaIN1d(style_dim, dim_in)
        self.norm2 = AdaIN1d(style_dim, dim_out)
        if self.learned_sc:
            self.conv1x1 = weight_norm(nn.Conv1d(dim_in, dim_out, 1, 1, 0, bias=False))
        if self.is_up:
            self.pool = weight_norm(nn.ConvTranspose1d(dim_in, dim_in, kernel_size=3, stride=2, groups=dim_in,
                                                       padding=1, output_padding=1))

    def build(self, dt, device):
        p = dict(conv1=nn_util.pack_conv1d(self.conv1, dt, device),
                 conv2=nn_util.pack_conv1d(self.conv2, dt, device),
                 sc=nn_util.pack_conv1d(self.conv1x1, dt, device) if self.learned_sc else None,
                 up_w=None, up_b=None)
        if self.is_up:
            w = nn_util.wn_weight(self.pool)                    # [C, 1, 3]
            p["up_w"] = w.reshape(-1, 3).contiguous().to(device)
            p["up_b"] = self.pool.bias.detach().float().contiguous().to(device)
        return p


class StyleFC:
    """All AdaIN ``fc`` layers of a module batched into ONE GEMM from the style vector.

    Each entry may read only a slice ``[lo, hi)`` of the style (the reference slices the 512-d
    style per branch, models.py:499,597-599): its weight is zero-padded to the full style width so
    a single ``[B, S] x [S, sum 2C]`` contraction produces every gamma/beta of the module."""

    def __init__(self, style_width: int):
        self.style_width = style_width
        self.entries = []   # (norm module, lo, hi)

    def add(self, norm: AdaIN1d, lo: int, hi: int) -> int:
        self.entries.append((norm, lo, hi))
        return len(self.entries) - 1

    def build(self, dt, device):
        ws, bs, offs = [], [], []
        off = 0
        for norm, lo, hi in self.entries:
            w = norm.fc.weight.detach().float()
            assert w.shape[1] == hi - lo, (w.shape, lo, hi)
            full = torch.zeros(w.shape[0], self.style_width)
            full[:, lo:hi] = w
            ws.append(full)
            bs.append(norm.fc.bias.detach().float())
            offs.append((off, w.shape[0]))
            off += w.shape[0]
        return dict(fc=nn_util.pack_linear(torch.cat(ws), torch.cat(bs), dt, device), offs=offs)

    @staticmethod
    def run(plan, style16: torch.Tensor):
        """``style16`` [B, S] 16-bit -> list of fp32 views [B, 2C] (gamma | beta) per entry."""
        B = style16.shape[0]
        out, _ = ops.conv(style16.view(1, B, -1), plan["fc"], raw=torch.float32)
        out = out.view(B, -1)
        return [out[:, o:o + n] for o, n in plan["offs"]]


def run_adain_block(p, x, gb1, gb2, lens, dt, out=None, out16=None, x16=None, res_dtype=None,
                    mid_dtype=torch.float32, lens_up=None):
    """AdainResBlk1d.forward (models.py:183-202) on channels-last ``x`` [B, T, Cin].

    ``x`` is the residual stream (fp32 or 16-bit): it feeds the InstanceNorm statistics, AdaIN and the
    identity shortcut.  A learned (1x1) shortcut needs a 16-bit GEMM operand: ``x16`` (defaults to
    ``x`` when that already is 16-bit).  The result is written as ``res_dtype`` (default ``dt``) into
    ``out`` (or a new tensor) and, when ``out16`` is given (dtype or view), additionally as a 16-bit
    copy for the next block's 1x1 shortcut.  Returns (out, out16, lens'); T' = 2T for upsample."""
    a1 = ops.adain_norm(x, gb1, LRELU, lens, dt, p["up_w"], p["up_b"])
    up = p["up_w"] is not None
    # ``lens_up`` = 2 * lens when the caller already holds it (saves a tiny elementwise launch per upsampling block)
    lens2 = (lens_up if lens_up is not None else lens * 2) if (up and lens is not None) else lens
    c1, _ = ops.conv(a1, p["conv1"], raw=mid_dtype, lens=lens2)
    a2 = ops.adain_norm(c1, gb2, LRELU, lens2, dt)
    if p["sc"] is not None:
        xs = x16 if x16 is not None else x
        if xs.dtype != dt:
            raise ValueError("run_adain_block: a learned shortcut needs the 16-bit copy of x (x16=)")
        if up:
            xs = ops.repeat_rows(xs, 2, lens, dt)
        xs, _ = ops.conv(xs, p["sc"], raw=torch.float32)
    else:
        xs = ops.repeat_rows(x, 2, lens) if up else x               # nearest x2 (models.py:261-270)
    raw_spec = out if out is not None else (res_dtype or dt)
    res, res16 = ops.conv(a2, p["conv2"], res1=xs, scale=INV_SQRT2, raw=raw_spec, act_out=out16, lens=lens2)
    return res, res16, lens2


# --------------------------------------------------------------------------------------------
# 2-D / 1-D residual blocks of the style stacks (models.py:19-156), no normalisation
# --------------------------------------------------------------------------------------------
_DOWN = {  # layer_type -> (depthwise (kh, kw), stride (sh, sw), padding (ph, pw)) in the reference's [H, W]
    "half": ((3, 3), (2, 2), (1, 1)),
    "channelpreserve": ((1, 3), (1, 2), (0, 1)),
    "timepreserve": ((3, 1), (2, 1), (1, 0)),
}


class LearnedDownSample(nn.Module):
    def __init__(self, layer_type: str, dim_in: int):
        super().__init__()
        self.layer_type = layer_type
        if layer_type == "none":
            self.conv = nn.Identity()
        else:
            k, s, p = _DOWN[layer_type]
            self.conv = spectral_norm(nn.Conv2d(dim_in, dim_in, kernel_size=k, stride=s, groups=dim_in, padding=p))


class ResBlk(nn.Module):
    """2-D residual block container (conv1, conv2, conv1x1, downsample_res.conv), models.py:59-100."""

    def __init__(self, dim_in, dim_out, downsample="none"):
        super().__init__()
        self.dim_in, self.dim_out, self.down = dim_in, dim_out, downsample
        self.learned_sc = dim_in != dim_out
        self.downsample_res = LearnedDownSample(downsample, dim_in)
        self.conv1 = spectral_norm(nn.Conv2d(dim_in, dim_in, 3, 1, 1))
        self.conv2 = spectral_norm(nn.Conv2d(dim_in, dim_out, 3, 1, 1))
        if self.learned_sc:
            self.conv1x1 = spectral_norm(nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=False))

    def build(self, dt, device):
        if self.down not in ("half", "channelpreserve"):
            raise NotImplementedError(f"ResBlk downsample={self.down!r} is not used by ArtSpeech")
        (kh, kw), (sh, sw), (ph, pw) = _DOWN[self.down]
        dw_w, dw_b = nn_util.dw_weight_2d(self.downsample_res.conv, device)
        return dict(conv1=nn_util.pack_conv2d(self.conv1, dt, device, (1, 1)),
                    conv2=nn_util.pack_conv2d(self.conv2, dt, device, (1, 1)),
                    sc=nn_util.pack_conv2d(self.conv1x1, dt, device, (0, 0)) if self.learned_sc else None,
                    dw_w=dw_w, dw_b=dw_b,
                    # (t, f) = (w, h)
                    dw_k=(kw, kh), dw_s=(sw, sh), dw_p=(pw, ph), pool=(sw, sh))


def run_resblk2d(p, x_raw, x_act, dt, want_raw=True):
    """ResBlk.forward: ``(avgpool(conv1x1(x)) + conv2(lrelu(dw(conv1(lrelu(x)))))) / sqrt(2)``.

    ``x_raw``/``x_act`` [B,T,F,C] (x and lrelu(x)).  The 1x1 shortcut conv is applied after the
    average pool (both linear, bias-free: they commute) to cut its work by the pooling factor."""
    c1, _ = ops.conv(x_act, p["conv1"], raw=dt)
    d = ops.dwconv(c1, p["dw_w"], p["dw_b"], p["dw_k"], p["dw_s"], p["dw_p"], act=ops.ACT_LRELU, slope=LRELU,
                   out_dtype=dt)
    sc = ops.avgpool(x_raw, p["pool"][0], p["pool"][1], dt)
    if p["sc"] is not None:
        sc, _ = ops.conv(sc, p["sc"], raw=torch.float32)
    return ops.conv(d, p["conv2"], res1=sc, scale=INV_SQRT2, raw=dt if want_raw else None, act_out=dt,
                    act=ops.ACT_LRELU, slope=LRELU)


class ResBlk1d(nn.Module):
    """1-D residual block container (weight-normed), models.py:102-156."""

    def __init__(self, dim_in, dim_out, downsample="none", dropout_p=0.2):
        super().__init__()
        self.dim_in, self.dim_out = dim_in, dim_out
        self.downsample_type = downsample
        self.learned_sc = dim_in != dim_out
        self.conv1 = weight_norm(nn.Conv1d(dim_in, dim_in, 3, 1, 1))
        self.conv2 = weight_norm(nn.Conv1d(dim_in, dim_out, 3, 1, 1))
        if self.learned_sc:
            self.conv1x1 = weight_norm(nn.Conv1d(dim_in, dim_out, 1, 1, 0, bias=False))
        if downsample == "none":
            self.pool = nn.Identity()
        else:
            self.pool = weight_norm(nn.Conv1d(dim_in, dim_in, kernel_size=3, stride=2, groups=dim_in, padding=1))

    def build(self, dt, device):
        if self.downsample_type is not True:
            raise NotImplementedError("ResBlk1d without downsample=True is not used by ArtSpeech")
        dw_w, dw_b = nn_util.dw_weight_1d(self.pool, device)
        return dict(conv1=nn_util.pack_conv1d(self.conv1, dt, device),
                    conv2=nn_util.pack_conv1d(self.conv2, dt, device),
                    sc=nn_util.pack_conv1d(self.conv1x1, dt, device) if self.learned_sc else None,
                    dw_w=dw_w, dw_b=dw_b)


def run_resblk1d(p, x_raw, x_act, dt, want_raw=True):
    """ResBlk1d.forward with ``downsample=True`` (models.py:127-156) on [B,T,C]."""
    c1, _ = ops.conv(x_act, p["conv1"], raw=dt)
    d = ops.dwconv(c1, p["dw_w"], p["dw_b"], (3, 1), (2, 1), (1, 0), act=ops.ACT_LRELU, slope=LRELU, out_dtype=dt)
    sc = ops.avgpool(x_raw, 2, 1, dt)
    if p["sc"] is not None:
        sc, _ = ops.conv(sc, p["sc"], raw=torch.float32)
    return ops.conv(d, p["conv2"], res1=sc, scale=INV_SQRT2, raw=dt if want_raw else None, act_out=dt,
                    act=ops.ACT_LRELU, slope=LRELU)


class Placeholder(nn.Identity):
    """Parameter-less slot that keeps nn.Sequential indices aligned with the reference."""
