"""Relative-position Transformer text encoder on the sm_100a kernels.

Drop-in for ``Utils/RelTransformerEnc.py``: ``RelTransformerEncoder(n_layers, hidden_channels, ...)``
keeps the constructor defaults, parameter names/shapes (``emb``, ``pre.conv_layers/norm_layers/proj``,
``encoder.attn_layers.{i}.conv_q/k/v/o + emb_rel_k/v``, ``norm_layers_1/2.{i}.gamma/beta``,
``ffn_layers.{i}.conv_1/conv_2``, ``last_ln``) and ``forward(x, x_lengths) -> [B, T, H]``
(RelTransformerEnc.py:328-380).

Dataflow per call (channels-last, fp32 residual stream, 16-bit GEMM operands):
  embed*sqrt(H) -> 3x [conv k5 -> channel-LN(eps 1e-4)+ReLU] -> 1x1 proj + residual (:316-325)
  -> n_layers x { LN -> fused q|k|v GEMM (N=3H) -> fused windowed rel-pos attention (:138-169)
                  -> conv_o + residual -> LN -> conv k9 + ReLU -> 1x1 + residual (:67-87) }
  -> last LN.  Every epilogue zeroes rows beyond ``x_lengths`` (the reference's ``* x_mask``).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import nn_util, ops

_pad = "$"
_punctuation = ';:,.!?¡¿—…"«»“” '
_letters = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
_letters_ipa = "ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞↓↑→↗↘'̩'ᵻ"
# 178 phoneme symbols, id 0 = pad (the table test.py and meldataset.py index into)
symbols = [_pad] + list(_punctuation) + list(_letters) + list(_letters_ipa)

LN_EPS = 1e-4


class LayerNorm(nn.Module):
    """Channel LayerNorm parameters (``gamma``/``beta``), RelTransformerEnc.py:272-290."""

    def __init__(self, channels: int, eps: float = LN_EPS):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


class MultiHeadAttention(nn.Module):
    def __init__(self, channels: int, out_channels: int, n_heads: int, window_size=None):
        super().__init__()
        assert channels % n_heads == 0
        self.channels, self.n_heads, self.window_size = channels, n_heads, window_size
        self.k_channels = channels // n_heads
        for name in ("conv_q", "conv_k", "conv_v"):
            setattr(self, name, nn.Conv1d(channels, channels, 1))
        if window_size is not None:
            std = self.k_channels ** -0.5
            self.emb_rel_k = nn.Parameter(torch.randn(1, 2 * window_size + 1, self.k_channels) * std)
            self.emb_rel_v = nn.Parameter(torch.randn(1, 2 * window_size + 1, self.k_channels) * std)
        self.conv_o = nn.Conv1d(channels, out_channels, 1)
        for c in (self.conv_q, self.conv_k, self.conv_v):
            nn.init.xavier_uniform_(c.weight)


class FFN(nn.Module):
    def __init__(self, in_channels, out_channels, filter_channels, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.conv_2 = nn.Conv1d(filter_channels, out_channels, 1)


class ConvReluNorm(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, kernel_size, n_layers):
        super().__init__()
        self.n_layers, self.kernel_size = n_layers, kernel_size
        chans = [in_channels] + [hidden_channels] * n_layers
        self.conv_layers = nn.ModuleList(
            nn.Conv1d(chans[i], chans[i + 1], kernel_size, padding=kernel_size // 2) for i in range(n_layers))
        self.norm_layers = nn.ModuleList(LayerNorm(hidden_channels) for _ in range(n_layers))
        self.proj = nn.Conv1d(hidden_channels, out_channels, 1)
        self.proj.weight.data.zero_()
        self.proj.bias.data.zero_()


class Encoder(nn.Module):
    def __init__(self, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, window_size, pre_ln):
        super().__init__()
        self.n_layers, self.pre_ln = n_layers, pre_ln
        self.attn_layers = nn.ModuleList(
            MultiHeadAttention(hidden_channels, hidden_channels, n_heads, window_size) for _ in range(n_layers))
        self.norm_layers_1 = nn.ModuleList(LayerNorm(hidden_channels) for _ in range(n_layers))
        self.ffn_layers = nn.ModuleList(
            FFN(hidden_channels, hidden_channels, filter_channels, kernel_size) for _ in range(n_layers))
        self.norm_layers_2 = nn.ModuleList(LayerNorm(hidden_channels) for _ in range(n_layers))
        if pre_ln:
            self.last_ln = LayerNorm(hidden_channels)


class RelTransformerEncoder(nn_util.PlanMixin, nn.Module):
    def __init__(self, n_layers, hidden_channels, kernel_size=9, n_heads=4, p_dropout=0.0, window_size=4,
                 block_length=None, prenet=True, pre_ln=True):
        super().__init__()
        if block_length is not None or not prenet or not pre_ln or window_size is None:
            raise NotImplementedError("only the configuration ArtSpeech instantiates is implemented "
                                      "(prenet, pre-LN, windowed relative attention, no block mask)")
        self.n_vocab = len(symbols)
        self.hidden_channels = hidden_channels
        self.filter_channels = hidden_channels * 2
        self.n_heads, self.n_layers = n_heads, n_layers
        self.kernel_size, self.window_size = kernel_size, window_size
        self.emb = nn.Embedding(self.n_vocab, hidden_channels, padding_idx=0)
        nn.init.normal_(self.emb.weight, mean=0, std=hidden_channels ** -0.5)
        nn.init.constant_(self.emb.weight[0], 0)
        self.pre = ConvReluNorm(hidden_channels, hidden_channels, hidden_channels, kernel_size=5, n_layers=3)
        self.encoder = Encoder(hidden_channels, self.filter_channels, n_heads, n_layers, kernel_size,
                               window_size, pre_ln)
        self.compute_dtype = torch.float16
        self._init_plan()

    def _build_plan(self, device):
        dt = self.compute_dtype
        f32 = lambda t: t.detach().float().contiguous().to(device)
        p = {"emb": f32(self.emb.weight)}
        p["pre_conv"] = [nn_util.pack_conv1d(c, dt, device) for c in self.pre.conv_layers]
        p["pre_ln"] = [(f32(n.gamma), f32(n.beta)) for n in self.pre.norm_layers]
        p["pre_proj"] = nn_util.pack_conv1d(self.pre.proj, dt, device)
        layers = []
        enc = self.encoder
        for i in range(self.n_layers):
            a = enc.attn_layers[i]
            wqkv = torch.cat([a.conv_q.weight, a.conv_k.weight, a.conv_v.weight], dim=0).squeeze(-1)
            bqkv = torch.cat([a.conv_q.bias, a.conv_k.bias, a.conv_v.bias], dim=0)
            layers.append(dict(
                ln1=(f32(enc.norm_layers_1[i].gamma), f32(enc.norm_layers_1[i].beta)),
                qkv=nn_util.pack_linear(wqkv, bqkv, dt, device),
                rel_k=f32(a.emb_rel_k[0]), rel_v=f32(a.emb_rel_v[0]),
                out=nn_util.pack_conv1d(a.conv_o, dt, device),
                ln2=(f32(enc.norm_layers_2[i].gamma), f32(enc.norm_layers_2[i].beta)),
                ffn1=nn_util.pack_conv1d(enc.ffn_layers[i].conv_1, dt, device),
                ffn2=nn_util.pack_conv1d(enc.ffn_layers[i].conv_2, dt, device),
            ))
        p["layers"] = layers
        p["last_ln"] = (f32(enc.last_ln.gamma), f32(enc.last_ln.beta))
        return p

    @torch.no_grad()
    def forward(self, x, x_lengths, want_16bit: bool = False):
        """``x`` int64 [B, T], ``x_lengths`` [B] -> fp32 [B, T, H] (zeros beyond each length).

        ``want_16bit`` additionally returns the same tensor in the compute dtype (operand of the
        next GEMM) so callers need no extra cast pass."""
        dev = x.device
        p = self.plan(dev)
        dt = self.compute_dtype
        lens = x_lengths.to(device=dev, dtype=torch.int32)
        H = self.hidden_channels
        x32, h16 = ops.embed(x, p["emb"], math.sqrt(H), lens, out32=True, out16=dt)
        for i in range(self.pre.n_layers):
            raw, _ = ops.conv(h16, p["pre_conv"][i], raw=torch.float32)
            g, b = p["pre_ln"][i]
            _, h16 = ops.layernorm(raw, g, b, LN_EPS, act=ops.ACT_RELU, lens=lens, out_b=dt)
        x32, _ = ops.conv(h16, p["pre_proj"], res1=x32, raw=torch.float32, lens=lens)
        for lp in p["layers"]:
            _, n16 = ops.layernorm(x32, *lp["ln1"], LN_EPS, lens=lens, out_b=dt)
            qkv, _ = ops.conv(n16, lp["qkv"], raw=torch.float32)
            a16 = ops.relpos_attention(qkv, lp["rel_k"], lp["rel_v"], self.window_size, self.n_heads, lens, dt)
            x32, _ = ops.conv(a16, lp["out"], res1=x32, raw=torch.float32, lens=lens)
            _, n16 = ops.layernorm(x32, *lp["ln2"], LN_EPS, lens=lens, out_b=dt)
            _, f16 = ops.conv(n16, lp["ffn1"], act_out=dt, act=ops.ACT_RELU, lens=lens)
            x32, _ = ops.conv(f16, lp["ffn2"], res1=x32, raw=torch.float32, lens=lens)
        out32, out16 = ops.layernorm(x32, *p["last_ln"], LN_EPS, lens=lens, out_a=torch.float32,
                                     out_b=dt if want_16bit else None)
        return (out32, out16) if want_16bit else out32
