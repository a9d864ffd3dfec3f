"""TEST INFRASTRUCTURE — plain-torch fp32 restatement of the reference's synthesis forward pass.

Each function restates one reference ``forward`` from a *state_dict* (no nn.Module, no reference
import), so it travels to the GPU box where ``/root/reference`` does not exist.  It is pinned
against the real reference modules by ``tests/test_oracle_pinning.py`` (runs where the reference is
mounted) and against the committed fixtures in ``tests/golden`` (generated from the reference by
``oracle/make_golden.py``).  It is the checker for the CUDA path and the timed ``cpu_baseline`` /
``--impl reference`` arm of ``bench.py`` — never part of the product path.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# parametrisations
# --------------------------------------------------------------------------------------------
def wn_weight(sd, prefix):
    """Old-style weight_norm (dim=0): ``g * v / ||v||``; plain ``weight`` once removed."""
    if prefix + ".weight_g" in sd:
        v, g = sd[prefix + ".weight_v"].float(), sd[prefix + ".weight_g"].float()
        n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / n)
    return sd[prefix + ".weight"].float()


def sn_weight(sd, prefix):
    """Old-style spectral_norm in eval mode: ``W / sigma``, ``sigma = u . (W_mat v)`` with the
    stored ``u, v`` (no power iteration)."""
    if prefix + ".weight_orig" in sd:
        w = sd[prefix + ".weight_orig"].float()
        u, v = sd[prefix + ".weight_u"].float(), sd[prefix + ".weight_v"].float()
        sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
        return w / sigma
    return sd[prefix + ".weight"].float()


def any_weight(sd, prefix):
    if prefix + ".weight_orig" in sd:
        return sn_weight(sd, prefix)
    return wn_weight(sd, prefix)


def opt(sd, key):
    return sd[key].float() if key in sd else None


# --------------------------------------------------------------------------------------------
# Vocoder/vocoder.py
# --------------------------------------------------------------------------------------------
VOCODER_CFG = dict(upsample_rates=(10, 5, 3, 2), upsample_kernel_sizes=(20, 10, 6, 4),
                   resblock_kernel_sizes=(3, 7, 11), resblock_dilation_sizes=((1, 3, 5),) * 3)


def _pad(k, d=1):
    return (k * d - d) // 2


def resblock1(sd, prefix, x, k, dilations):
    """ResBlock1.forward, vocoder.py:35-42."""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, wn_weight(sd, f"{prefix}.convs1.{m}"), opt(sd, f"{prefix}.convs1.{m}.bias"),
                      padding=_pad(k, d), dilation=d)
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, wn_weight(sd, f"{prefix}.convs2.{m}"), opt(sd, f"{prefix}.convs2.{m}.bias"),
                      padding=_pad(k, 1))
        x = xt + x
    return x


@torch.no_grad()
def generator_forward(sd, mel, cfg=VOCODER_CFG):
    """Generator.forward, vocoder.py:100-116.  ``mel`` [B, 80, T] fp32 -> wav [B, 1, 300 T]."""
    x = F.conv1d(mel.float(), wn_weight(sd, "conv_pre"), opt(sd, "conv_pre.bias"), padding=3)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, wn_weight(sd, f"ups.{i}"), opt(sd, f"ups.{i}.bias"), stride=u,
                               padding=u // 2 + u % 2, output_padding=u % 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = resblock1(sd, f"resblocks.{i * nk + j}", x, rk, rd)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (vocoder.py:112)
    x = F.conv1d(x, wn_weight(sd, "conv_post"), opt(sd, "conv_post.bias"), padding=3)
    return torch.tanh(x)
