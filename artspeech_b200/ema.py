"""EMA (vocal-tract) extractor on the sm_100a kernels — drop-in for
``Utils/EMA/EMA_Predictor.py`` + the vendored conformer blocks
(``Utils/EMA/conformer/conformer/{encoder,attention,convolution,feed_forward}.py``).

``EMA_Predictor().forward(F0[B,1,T], energy[B,1,T], mels[B,80,T]) -> [B,10,T]`` (:65-82).
Parameter tree identical to the reference (``encoder1``, ``decoder.{i}.sequential.{0..4}...``,
``pool``, ``decoder2``, ``decoder3``) so ``200000.pth.tar`` loads unchanged.

Reference quirks reproduced on purpose:
  * ``decoder2`` is an nn.LSTM without ``batch_first`` fed [B,T,256]: at batch 1 it is ONE step
    from zero state per frame, so ``W_hh`` and the forget gate are dead (as_lstm_onestep);
  * attention scores are scaled by sqrt(d_model)=16, not sqrt(d_head), and use the view-based
    relative shift on absolute sinusoids (attention.py:57,91,105-113).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm

from . import nn_util, ops
from .blocks import Placeholder

LN_EPS = 1e-5


class _Linear(nn.Module):
    """``modules.Linear`` wrapper: parameters live under ``.linear`` (modules.py:38-52)."""

    def __init__(self, i, o, bias=True):
        super().__init__()
        self.linear = nn.Linear(i, o, bias=bias)
        nn.init.xavier_uniform_(self.linear.weight)
        if bias:
            nn.init.zeros_(self.linear.bias)


class _Conv(nn.Module):
    """Pointwise / depthwise wrappers keep the conv under ``.conv`` (convolution.py:48-57,93-102)."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class _Slot(nn.Module):
    """``ResidualConnectionModule``: the wrapped module lives under ``.module`` (modules.py:22-35)."""

    def __init__(self, module, factor=1.0):
        super().__init__()
        self.module = module
        self.module_factor = factor


class _FeedForward(nn.Module):
    def __init__(self, d, expansion=4):
        super().__init__()
        self.sequential = nn.Sequential(nn.LayerNorm(d), _Linear(d, d * expansion), Placeholder(), Placeholder(),
                                        _Linear(d * expansion, d), Placeholder())


class _PositionalEncoding(nn.Module):
    def __init__(self, d_model=512, max_len=10000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _RelAttention(nn.Module):
    def __init__(self, d_model, num_heads):
        super().__init__()
        self.num_heads, self.d_head = num_heads, d_model // num_heads
        self.query_proj = _Linear(d_model, d_model)
        self.key_proj = _Linear(d_model, d_model)
        self.value_proj = _Linear(d_model, d_model)
        self.pos_proj = _Linear(d_model, d_model, bias=False)
        self.u_bias = nn.Parameter(torch.empty(num_heads, self.d_head))
        self.v_bias = nn.Parameter(torch.empty(num_heads, self.d_head))
        nn.init.xavier_uniform_(self.u_bias)
        nn.init.xavier_uniform_(self.v_bias)
        self.out_proj = _Linear(d_model, d_model)


class _MHSA(nn.Module):
    def __init__(self, d_model, num_heads):
        super().__init__()
        self.positional_encoding = _PositionalEncoding(d_model)
        self.layer_norm = nn.LayerNorm(d_model)
        self.attention = _RelAttention(d_model, num_heads)


class _ConvModule(nn.Module):
    def __init__(self, d, kernel_size=31):
        super().__init__()
        self.kernel_size = kernel_size
        self.sequential = nn.Sequential(
            nn.LayerNorm(d), Placeholder(),
            _Conv(nn.Conv1d(d, 2 * d, 1)), Placeholder(),
            _Conv(nn.Conv1d(d, d, kernel_size, groups=d, padding=(kernel_size - 1) // 2, bias=False)),
            nn.BatchNorm1d(d), Placeholder(), _Conv(nn.Conv1d(d, d, 1)), Placeholder())


class ConformerBlock(nn.Module):
    def __init__(self, encoder_dim=256, num_attention_heads=4, conv_kernel_size=31):
        super().__init__()
        self.sequential = nn.Sequential(
            _Slot(_FeedForward(encoder_dim), 0.5),
            _Slot(_MHSA(encoder_dim, num_attention_heads)),
            _Slot(_ConvModule(encoder_dim, conv_kernel_size)),
            _Slot(_FeedForward(encoder_dim), 0.5),
            nn.LayerNorm(encoder_dim))


class EMA_Predictor(nn_util.PlanMixin, nn.Module):
    D = 256
    HEADS = 4

    def __init__(self):
        super().__init__()
        self.encoder1 = nn.Sequential(nn.Linear(82, 256), Placeholder(), nn.BatchNorm1d(256), Placeholder(),
                                      Placeholder(), Placeholder())
        self.decoder = nn.ModuleList([ConformerBlock(256, 4, 31) for _ in range(3)])
        self.pool = weight_norm(nn.ConvTranspose1d(256, 256, kernel_size=3, stride=2, groups=256, padding=1,
                                                   output_padding=1))   # constructed, never used (:43)
        self.decoder2 = nn.LSTM(input_size=256, hidden_size=256, num_layers=1, dropout=0, bidirectional=True)
        self.decoder3 = nn.Sequential(nn.Linear(512, 128), Placeholder(), nn.BatchNorm1d(128), Placeholder(),
                                      Placeholder(), nn.Linear(128, 10))
        self.compute_dtype = torch.float16
        self._init_plan()

    def _build_plan(self, device):
        dt = self.compute_dtype
        f32 = lambda t: t.detach().float().contiguous().to(device)
        lin = lambda l, **kw: nn_util.pack_linear(l.weight, l.bias, dt, device, **kw)
        p = {}
        s, sh = nn_util.bn_affine(self.encoder1[2])
        p["enc1"] = lin(self.encoder1[0], scale=s, shift=sh)
        blocks = []
        for blk in self.decoder:
            seq = blk.sequential
            bp = {}
            for name, slot in (("ff1", seq[0]), ("ff2", seq[3])):
                ff = slot.module.sequential
                bp[name] = dict(ln=(f32(ff[0].weight), f32(ff[0].bias)), w1=lin(ff[1].linear), w2=lin(ff[4].linear))
            mh = seq[1].module
            at = mh.attention
            wqkv = torch.cat([at.query_proj.linear.weight, at.key_proj.linear.weight, at.value_proj.linear.weight])
            bqkv = torch.cat([at.query_proj.linear.bias, at.key_proj.linear.bias, at.value_proj.linear.bias])
            bp["att"] = dict(ln=(f32(mh.layer_norm.weight), f32(mh.layer_norm.bias)),
                             qkv=nn_util.pack_linear(wqkv, bqkv, dt, device),
                             # pos_proj(PE[0:T]) does not depend on the input: fold PE @ W^T lazily per T
                             pos_w=f32(at.pos_proj.linear.weight), pe=f32(mh.positional_encoding.pe[0]),
                             pos_cache={},
                             u=f32(at.u_bias), v=f32(at.v_bias), out=lin(at.out_proj.linear))
            cm = seq[2].module.sequential
            s, sh = nn_util.bn_affine(cm[5])
            dw_w, dw_b = nn_util.dw_weight_1d(cm[4].conv, device, s, sh)
            bp["conv"] = dict(ln=(f32(cm[0].weight), f32(cm[0].bias)),
                              pw1=nn_util.pack_conv1d(cm[2].conv, dt, device), dw_w=dw_w, dw_b=dw_b,
                              k=cm[4].conv.kernel_size[0], pw2=nn_util.pack_conv1d(cm[7].conv, dt, device))
            bp["ln"] = (f32(seq[4].weight), f32(seq[4].bias))
            blocks.append(bp)
        p["blocks"] = blocks
        p["lstm_proj"], _ = nn_util.pack_lstm(self.decoder2, dt, device)
        s, sh = nn_util.bn_affine(self.decoder3[2])
        p["dec3a"] = lin(self.decoder3[0], scale=s, shift=sh)
        p["dec3b"] = lin(self.decoder3[5])
        return p

    @staticmethod
    def _pos(ap, T):
        """pos_proj(PE[0:T]) as fp32 [T, 256]; input-independent, cached per T (tiny fp32 GEMM done
        once per length on the host-side plan, exactly PE @ W^T)."""
        if T not in ap["pos_cache"]:
            ap["pos_cache"][T] = (ap["pe"][:T] @ ap["pos_w"].t()).contiguous()
        return ap["pos_cache"][T]

    @torch.no_grad()
    def forward_cl(self, feat16: torch.Tensor, lens=None) -> torch.Tensor:
        """``feat16`` [B, T, 82] view (F0 | energy | mel), 16-bit -> EMA fp32 [B, T, 10]."""
        p = self.plan(feat16.device)
        dt = self.compute_dtype
        B, T, _ = feat16.shape
        D, H = self.D, self.HEADS
        _, x32 = ops.conv(feat16, p["enc1"], act_out=torch.float32, act=ops.ACT_RELU)   # Linear+BN+ReLU
        x16 = None
        for bp in p["blocks"]:
            x32 = self._ff(bp["ff1"], x32, dt)
            ap = bp["att"]
            _, n16 = ops.layernorm(x32, *ap["ln"], LN_EPS, out_b=dt)
            qkv, _ = ops.conv(n16, ap["qkv"], raw=torch.float32)
            a16 = ops.conformer_attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], self._pos(ap, T),
                                          ap["u"], ap["v"], H, lens, dt)
            x32, _ = ops.conv(a16, ap["out"], res1=x32, raw=torch.float32)
            cp = bp["conv"]
            _, n16 = ops.layernorm(x32, *cp["ln"], LN_EPS, out_b=dt)
            g, _ = ops.conv(n16, cp["pw1"], raw=dt)                                   # [B,T,512] for the GLU
            k = cp["k"]
            d16 = ops.dwconv(g, cp["dw_w"], cp["dw_b"], (k, 1), (1, 1), ((k - 1) // 2, 0), glu=True,
                             act=ops.ACT_SWISH, out_dtype=dt, lens_in=lens)
            x32, _ = ops.conv(d16, cp["pw2"], res1=x32, raw=torch.float32)
            x32 = self._ff(bp["ff2"], x32, dt)
            x32, x16 = ops.layernorm(x32, *bp["ln"], LN_EPS, out_a=torch.float32, out_b=dt)
        xproj, _ = ops.conv(x16, p["lstm_proj"], raw=torch.float32)
        h16 = ops.lstm_onestep(xproj, 256, dt)
        _, z16 = ops.conv(h16, p["dec3a"], act_out=dt, act=ops.ACT_RELU)
        ema, _ = ops.conv(z16, p["dec3b"], raw=torch.float32)
        return ema

    @staticmethod
    def _ff(fp, x32, dt):
        # x + 0.5 * W2(swish(W1(LN(x))))   (feed_forward.py:47-57, encoder.py:74-83)
        _, n16 = ops.layernorm(x32, *fp["ln"], LN_EPS, out_b=dt)
        _, h16 = ops.conv(n16, fp["w1"], act_out=dt, act=ops.ACT_SWISH)
        # (W2 h + b2) * 0.5 + x  ==  (W2 h + b2 + 2x) * 0.5
        y, _ = ops.conv(h16, fp["w2"], res1=x32, res2=x32, scale=0.5, raw=torch.float32)
        return y

    @torch.no_grad()
    def forward(self, F0, energy, mels=None):
        """Reference signature: [B,1,T], [B,1,T], [B,80,T] -> [B,10,T]."""
        B, _, T = F0.shape
        dt = self.compute_dtype
        feat = torch.zeros(B, T, 88, dtype=dt, device=F0.device)
        ops.to_channels_last(F0.float(), dt, out=feat[..., 0:1])
        ops.to_channels_last(energy.float(), dt, out=feat[..., 1:2])
        ops.to_channels_last(mels.float(), dt, out=feat[..., 2:82])
        ema = self.forward_cl(feat[..., :82])
        return ops.to_channels_first(ema, torch.float32)
