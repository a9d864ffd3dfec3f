"""JDCNet F0 extractor on the sm_100a kernels (drop-in for ``Utils/JDC/model.py``).

``JDCNet(num_class=1, seq_len=192).forward(mel[B,1,80,T]) -> |F0| [B,1,T]`` (model.py:102-137).
Same parameters/buffers as the reference, including the unused ``bilstm_detector`` / ``detector`` /
``detector_conv`` heads, so ``bst.t7`` loads unchanged.

Image layout: the reference transposes to [B,1,T,80]; here the image is channels-last
[B, T, F=80, C].  Every Conv2d 3x3 is an implicit GEMM; eval-mode BatchNorm that follows a conv is
folded into its weights, BatchNorm that precedes one (ResBlock.pre_conv, pool_block) is fused with
LeakyReLU(0.01) and the (1,k) max-pool in one kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import nn_util, ops
from .blocks import Placeholder

SLOPE = 0.01


class ResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.pre_conv = nn.Sequential(nn.BatchNorm2d(in_channels), Placeholder(), Placeholder())
        self.conv = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False),
                                  nn.BatchNorm2d(out_channels), Placeholder(),
                                  nn.Conv2d(out_channels, out_channels, 3, padding=1, bias=False))
        self.conv1by1 = nn.Conv2d(in_channels, out_channels, 1, bias=False) if in_channels != out_channels else None


class JDCNet(nn_util.PlanMixin, nn.Module):
    def __init__(self, num_class=722, seq_len=31, leaky_relu_slope=0.01):
        super().__init__()
        self.num_class = num_class
        self.conv_block = nn.Sequential(nn.Conv2d(1, 64, 3, padding=1, bias=False), nn.BatchNorm2d(64),
                                        Placeholder(), nn.Conv2d(64, 64, 3, padding=1, bias=False))
        self.res_block1 = ResBlock(64, 128)
        self.res_block2 = ResBlock(128, 192)
        self.res_block3 = ResBlock(192, 256)
        self.pool_block = nn.Sequential(nn.BatchNorm2d(256), Placeholder(), Placeholder(), Placeholder())
        self.detector_conv = nn.Sequential(nn.Conv2d(640, 256, 1, bias=False), nn.BatchNorm2d(256),
                                           Placeholder(), Placeholder())
        self.bilstm_classifier = nn.LSTM(input_size=512, hidden_size=256, batch_first=True, bidirectional=True)
        self.bilstm_detector = nn.LSTM(input_size=512, hidden_size=256, batch_first=True, bidirectional=True)
        self.classifier = nn.Linear(512, num_class)
        self.detector = nn.Linear(512, 2)
        self.compute_dtype = torch.float16
        self._init_plan()

    def _build_plan(self, device):
        dt = self.compute_dtype
        if self.num_class != 1:
            raise NotImplementedError("only the num_class=1 F0 regressor ArtSpeech loads is implemented")
        p = {}
        s, sh = nn_util.bn_affine(self.conv_block[1])
        p["c0"] = nn_util.small_conv2d(self.conv_block[0], device, (1, 1), s, sh, t_is_h=True)     # conv + BN folded
        p["c1"] = nn_util.pack_conv2d(self.conv_block[3], dt, device, (1, 1), t_is_h=True)
        p["res"] = []
        for blk in (self.res_block1, self.res_block2, self.res_block3):
            s0, sh0 = nn_util.bn_affine(blk.pre_conv[0])
            s1, sh1 = nn_util.bn_affine(blk.conv[1])
            p["res"].append(dict(
                pre=(s0.to(device), sh0.to(device)),
                conv_a=nn_util.pack_conv2d(blk.conv[0], dt, device, (1, 1), s1, sh1, t_is_h=True),
                conv_b=nn_util.pack_conv2d(blk.conv[3], dt, device, (1, 1), t_is_h=True),
                sc=nn_util.pack_conv2d(blk.conv1by1, dt, device, (0, 0))))
        s, sh = nn_util.bn_affine(self.pool_block[0])
        p["pool"] = (s.to(device), sh.to(device))
        # classifier input index in the reference is c*2 + f ([B,256,T,2] -> permute -> view 512);
        # ours is f*256 + c, so permute the LSTM's input columns once.
        perm = torch.tensor([c * 2 + f for f in range(2) for c in range(256)])
        p["lstm_proj"], p["whh_t"] = nn_util.pack_lstm(self.bilstm_classifier, dt, device, in_perm=perm)
        p["cls"] = nn_util.pack_linear(self.classifier.weight, self.classifier.bias, dt, device)
        return p

    @torch.no_grad()
    def forward_cl(self, img: torch.Tensor, lens=None) -> torch.Tensor:
        """``img`` [B, T, 80, 1] (any dtype) -> |F0| fp32 [B, T, 1]."""
        p = self.plan(img.device)
        dt = self.compute_dtype
        B, T = img.shape[0], img.shape[1]
        _, a = ops.conv_small(img, p["c0"], act_out=dt, act=ops.ACT_LRELU, slope=SLOPE)
        x, _ = ops.conv(a, p["c1"], raw=dt)                                         # convblock_out
        for rp in p["res"]:
            xp = ops.affine_act_maxpool(x, rp["pre"][0], rp["pre"][1], SLOPE, 2, dt)   # BN -> LReLU -> MaxPool(1,2)
            sc, _ = ops.conv(xp, rp["sc"], raw=torch.float32)                        # conv1by1(x)
            _, h = ops.conv(xp, rp["conv_a"], act_out=dt, act=ops.ACT_LRELU, slope=SLOPE)
            x, _ = ops.conv(h, rp["conv_b"], res1=sc, raw=dt)
        pooled = ops.affine_act_maxpool(x, p["pool"][0], p["pool"][1], SLOPE, 4, dt)  # [B,T,2,256]
        seq = pooled.view(B, T, 512)
        xproj, _ = ops.conv(seq, p["lstm_proj"], raw=torch.float32)
        h = ops.bilstm(xproj, p["whh_t"], 256, lens, dt)
        _, f0 = ops.conv(h, p["cls"], act_out=torch.float32, act=ops.ACT_ABS)
        return f0

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference signature: ``x`` [B,1,80,T] -> [B,1,T]."""
        B, _, M, T = x.shape
        img = ops.to_channels_last(x.reshape(B, M, T).float(), self.compute_dtype).view(B, T, M, 1)
        f0 = self.forward_cl(img)
        return f0.view(B, 1, T)
