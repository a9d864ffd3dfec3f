"""Weight preparation shared by the module classes: parametrisation folding and kernel packing.

The modules keep the reference's parameters untouched (so ``state_dict`` round-trips) and build a
*plan* — folded, re-laid-out, device-resident copies in the kernels' formats — lazily on the first
forward; ``load_state_dict`` / ``.to()`` invalidate it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


def wn_weight(m: nn.Module) -> torch.Tensor:
    """Old-style ``weight_norm`` (dim 0): ``g * v / ||v||``; plain ``weight`` when not wrapped."""
    if hasattr(m, "weight_g") and hasattr(m, "weight_v") and not hasattr(m, "weight_orig"):
        v = m.weight_v.detach().float()
        g = m.weight_g.detach().float()
        n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / n)
    return sn_weight(m)


def sn_weight(m: nn.Module) -> torch.Tensor:
    """Old-style ``spectral_norm`` in eval mode: ``weight_orig / (u . W_mat v)`` with the stored
    ``u, v`` — no power iteration."""
    if hasattr(m, "weight_orig"):
        w = m.weight_orig.detach().float()
        u, v = m.weight_u.detach().float(), m.weight_v.detach().float()
        sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
        return w / sigma
    return m.weight.detach().float()


def eff_weight(m: nn.Module) -> torch.Tensor:
    return sn_weight(m) if hasattr(m, "weight_orig") else wn_weight(m)


def bn_affine(bn: nn.Module):
    """Eval-mode BatchNorm as ``y = x * scale + shift`` (fp32)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def fbias(m: nn.Module) -> Optional[torch.Tensor]:
    b = getattr(m, "bias", None)
    return None if b is None else b.detach().float()


# ---- packing helpers -----------------------------------------------------------------------
def pack_conv1d(m: nn.Module, dtype, device, dilation: int = 1, scale=None, shift=None) -> ops.PackedConv:
    """Conv1d weight [Cout, Cin, k] -> implicit-GEMM taps.  Optional folded BatchNorm (scale, shift)
    applied to the OUTPUT channels."""
    w = eff_weight(m)                          # [Cout, Cin, k]
    b = fbias(m)
    if scale is not None:
        w = w * scale.reshape(-1, 1, 1)
        b = shift if b is None else b * scale + shift
    k = w.shape[2]
    return ops.pack_conv(w.permute(2, 0, 1), b, ops.taps_1d(k, dilation), dtype, device)


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor], dtype, device, scale=None,
                shift=None) -> ops.PackedConv:
    """Linear weight [Cout, Cin] as a 1-tap contraction."""
    w = weight.detach().float()
    b = None if bias is None else bias.detach().float()
    if scale is not None:
        w = w * scale.reshape(-1, 1)
        b = shift if b is None else b * scale + shift
    return ops.pack_conv(w.unsqueeze(0), b, [(0, 0)], dtype, device)


def _taps_hw(kh, kw, ph, pw, t_is_h):
    """Tap (dt, df) of kernel element (jh, jw).  The reference's images are [B, C, H, W]; ours are
    [B, T, F, C] with (T, F) = (W, H) for the style stacks (H = mel bins / EMA channels, W = time)
    and (T, F) = (H, W) for JDCNet, which transposes its input to [B, 1, T, 80]."""
    if t_is_h:
        return [(jh - ph, jw - pw) for jh in range(kh) for jw in range(kw)]
    return [(jw - pw, jh - ph) for jh in range(kh) for jw in range(kw)]


def pack_conv2d(m: nn.Module, dtype, device, pad_hw, scale=None, shift=None, t_is_h=False) -> ops.PackedConv:
    """Conv2d weight [Cout, Cin, kh, kw] -> implicit-GEMM taps over the channels-last image."""
    w = eff_weight(m)                          # [Cout, Cin, kh, kw]
    b = fbias(m)
    if scale is not None:
        w = w * scale.reshape(-1, 1, 1, 1)
        b = shift if b is None else b * scale + shift
    cout, cin, kh, kw = w.shape
    taps = _taps_hw(kh, kw, pad_hw[0], pad_hw[1], t_is_h)
    wt = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    return ops.pack_conv(wt, b, taps, dtype, device)


def small_conv2d(m: nn.Module, device, pad_hw, scale=None, shift=None, t_is_h=False) -> ops.SmallConv:
    """Same tap convention as ``pack_conv2d`` for the direct (tiny Cin) kernel, fp32 weights."""
    w = eff_weight(m)
    b = fbias(m)
    if scale is not None:
        w = w * scale.reshape(-1, 1, 1, 1)
        b = shift if b is None else b * scale + shift
    cout, cin, kh, kw = w.shape
    taps = _taps_hw(kh, kw, pad_hw[0], pad_hw[1], t_is_h)
    wt = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    return ops.pack_small_conv(wt, b, taps, device)


def small_conv1d(m: nn.Module, device) -> ops.SmallConv:
    w = eff_weight(m)                          # [Cout, Cin, k]
    k = w.shape[2]
    return ops.pack_small_conv(w.permute(2, 0, 1), fbias(m), ops.taps_1d(k), device)


def dw_weight_2d(m: nn.Module, device):
    """Depthwise Conv2d weight [C,1,kh,kw] -> fp32 [kt*kf, C] with (kt, kf) = (kw, kh)."""
    w = eff_weight(m)
    c, _, kh, kw = w.shape
    wt = w.reshape(c, kh, kw).permute(2, 1, 0).reshape(kw * kh, c).contiguous()
    return wt.to(device), None if fbias(m) is None else fbias(m).to(device)


def dw_weight_1d(m: nn.Module, device, scale=None, shift=None):
    """Depthwise Conv1d / ConvTranspose1d weight [C,1,k] -> fp32 [k, C] (optionally with a folded
    BatchNorm on the output channels)."""
    w = eff_weight(m)
    w = w.reshape(w.shape[0], -1)              # [C, k]
    b = fbias(m)
    if scale is not None:
        w = w * scale.reshape(-1, 1)
        b = shift if b is None else b * scale + shift
    return w.t().contiguous().to(device), None if b is None else b.contiguous().to(device)


def pack_lstm(lstm: nn.LSTM, dtype, device, in_perm: Optional[torch.Tensor] = None):
    """Bidirectional single-layer nn.LSTM -> (input projection PackedConv with N = 8H and bias
    b_ih + b_hh for forward|reverse, whh_t fp32 [2, H, 4H])."""
    wih = [lstm.weight_ih_l0.detach().float(), lstm.weight_ih_l0_reverse.detach().float()]
    whh = [lstm.weight_hh_l0.detach().float(), lstm.weight_hh_l0_reverse.detach().float()]
    bias = [(lstm.bias_ih_l0 + lstm.bias_hh_l0).detach().float(),
            (lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse).detach().float()]
    w = torch.cat(wih, dim=0)                  # [8H, In]
    if in_perm is not None:
        w = w[:, in_perm]
    proj = pack_linear(w, torch.cat(bias), dtype, device)
    whh_t = torch.stack([m.t().contiguous() for m in whh]).contiguous().to(device)
    return proj, whh_t


# Bumped whenever any module's plan is dropped.  CUDA graphs bake the addresses of the packed weights in, so the
# engine compares this counter before replaying and re-captures when weights were re-packed meanwhile.
_PLAN_EPOCH = [0]


def plan_epoch() -> int:
    return _PLAN_EPOCH[0]


def bump_plan_epoch() -> None:
    _PLAN_EPOCH[0] += 1


def param_signature(module: nn.Module):
    """Device and dtype of the module's parameters / buffers: ``.to()`` onto the device they already are on keeps
    it.  (Addresses are deliberately not part of it: ``nn.LSTM._apply`` re-flattens its weights into a fresh cuDNN
    buffer on every ``.to()``, while the plans hold their own packed copies.)"""
    return tuple((t.dtype, str(t.device)) for t in list(module.parameters()) + list(module.buffers()))


class PlanMixin:
    """Lazy plan cache for a module tree; invalidated by load_state_dict and by device / dtype moves that actually
    move the parameters (``model.to(device)`` on a model that already lives there keeps the packed weights, which
    captured CUDA graphs point at)."""

    def _init_plan(self):
        self._plan = None
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.invalidate_plan())

    def invalidate_plan(self):
        if self._plan is not None:
            bump_plan_epoch()
        self._plan = None
        for m in self.children():
            if isinstance(m, PlanMixin):
                m.invalidate_plan()

    def _apply(self, fn, *a, **kw):
        before = param_signature(self)
        out = super()._apply(fn, *a, **kw)
        if param_signature(self) != before and self._plan is not None:
            self._plan = None
            bump_plan_epoch()
        return out

    def plan(self, device):
        if self._plan is None or self._plan.get("_device") != str(device):
            if self._plan is not None:
                bump_plan_epoch()
            self._plan = self._build_plan(device)
            self._plan["_device"] = str(device)
        return self._plan
