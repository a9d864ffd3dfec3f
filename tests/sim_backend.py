"""Torch (CPU) model of the C-ABI primitives' *semantics*, for host-logic tests without a GPU.

``install(monkeypatch)`` swaps the functions in ``artspeech_b200.ops`` for these models so the
module classes (weight folding, polyphase packing, tap tables, dataflow) can be checked against the
oracle on CPU.  This lives under tests/ on purpose: the product has exactly one backend (CUDA) and
raises without it.  Each model follows the contract written in include/artspeech_b200.h.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from artspeech_b200 import ops

_ACT = {
    ops.ACT_NONE: lambda v, s: v,
    ops.ACT_LRELU: lambda v, s: torch.where(v > 0, v, v * s),
    ops.ACT_RELU: lambda v, s: torch.relu(v),
    ops.ACT_TANH: lambda v, s: torch.tanh(v),
    ops.ACT_SWISH: lambda v, s: v * torch.sigmoid(v),
    ops.ACT_ABS: lambda v, s: v.abs(),
}


def _write(spec, shape, value):
    if spec is None:
        return None
    if isinstance(spec, torch.dtype):
        return value.to(spec)
    spec.copy_(value.to(spec.dtype))
    return spec


def conv(x, pw, *, out_shape=None, use_bias=True, res1=None, res2=None, scale=1.0, raw=None,
         act_out=None, act=ops.ACT_NONE, slope=0.0, lens=None):
    squeeze = x.dim() == 3
    x4 = x.unsqueeze(2) if squeeze else x
    B, T, Fd, Cin = x4.shape
    assert Cin == pw.Cin and x.dtype == pw.w.dtype
    To, Fo = (T, Fd) if out_shape is None else out_shape
    xf = x4.float()
    w = pw.w.float()[:, :pw.Cout, :pw.Cin]          # [ntaps, Cout, Cin], 16-bit-rounded values
    acc = torch.zeros(B, To, Fo, pw.Cout)
    for j, (dt, df) in enumerate(pw.taps):
        # gather x[b, to+dt, fo+df, :] with zero fill
        t_idx = torch.arange(To) + dt
        f_idx = torch.arange(Fo) + df
        tv = (t_idx >= 0) & (t_idx < T)
        fv = (f_idx >= 0) & (f_idx < Fd)
        g = xf[:, t_idx.clamp(0, T - 1)][:, :, f_idx.clamp(0, Fd - 1)]
        g = g * (tv[None, :, None, None] & fv[None, None, :, None])
        acc += g @ w[j].t()
    if use_bias and pw.bias is not None:
        acc = acc + pw.bias.float()
    oshape = (B, To, pw.Cout) if squeeze else (B, To, Fo, pw.Cout)
    for r in (res1, res2):
        if r is not None:
            acc = acc + r.float().reshape(B, To, Fo, pw.Cout)
    acc = acc * scale
    if lens is not None:
        m = torch.arange(To)[None, :] < lens.long()[:, None]
        acc = acc * m[:, :, None, None]
    acc = acc.reshape(oshape)
    return _write(raw, oshape, acc), _write(act_out, oshape, _ACT[act](acc, slope))



def resblock_pair(xa, c1, c2, k, dil, *, slope, res2=None, res3=None, scale=1.0, out_act=ops.ACT_NONE,
                  out_slope=1.0, lens=None, out=None):
    """Contract of as_hifigan_resblock_pair (include/artspeech_b200.h): the stream is carried activated."""
    dt = xa.dtype
    a = xa.float()
    x_raw = torch.where(a < 0, a * (1.0 / slope), a)
    _, t = conv(xa, c1, act_out=dt, act=ops.ACT_LRELU, slope=slope, lens=lens)       # 16-bit operand of conv2
    v, _ = conv(t, c2, raw=torch.float32)
    v = v + x_raw
    for r in (res2, res3):
        if r is not None:
            v = v + r.float()
    v = v * scale
    if lens is not None:
        v = v * (torch.arange(v.shape[1])[None, :] < lens.long()[:, None])[:, :, None]
    y = _ACT[out_act](v, out_slope).to(dt)
    if out is not None:
        out.copy_(y)
        return out
    return y


def _mask_rows(y, lens, T):
    if lens is None:
        return y
    m = torch.arange(T)[None, :] < lens.long()[:, None]
    shape = [y.shape[0], T] + [1] * (y.dim() - 2)
    return y * m.reshape(shape)


def embed(tokens, table, scale, lens, *, out32, out16=None):
    v = table.float()[tokens] * scale
    v = _mask_rows(v, lens, tokens.shape[1])
    return (v.clone() if out32 else None), (v.to(out16) if out16 is not None else None)


def layernorm(x, gamma, beta, eps, *, act=ops.ACT_NONE, slope=0.0, lens=None, out_a=None, out_b=None):
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = ((xf - mean) ** 2).mean(-1, keepdim=True)
    y = (xf - mean) * torch.rsqrt(var + eps) * gamma.float() + beta.float()
    y = _mask_rows(_ACT[act](y, slope), lens, x.shape[1])
    return _write(out_a, y.shape, y), _write(out_b, y.shape, y)


def relpos_attention(qkv, emb_rel_k, emb_rel_v, window, n_heads, lens, out_dtype):
    B, T, C3 = qkv.shape
    HD = C3 // 3
    D = HD // n_heads
    q, k, v = [t.reshape(B, T, n_heads, D).transpose(1, 2) for t in qkv.float().split(HD, dim=-1)]
    ek, ev = emb_rel_k.float().reshape(-1, D), emb_rel_v.float().reshape(-1, D)
    out = torch.zeros(B, n_heads, T, D)
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]                     # j - i
    inwin = rel.abs() <= window
    relc = (rel + window).clamp(0, 2 * window)
    for b in range(B):
        L = T if lens is None else int(lens[b])
        if L == 0:
            continue
        s = q[b, :, :L] @ k[b, :, :L].transpose(-1, -2)
        qe = q[b, :, :L] @ ek.t()                          # [H, L, 2w+1]
        s = s + torch.where(inwin[:L, :L], torch.gather(qe, 2, relc[:L, :L].expand(n_heads, L, L)), torch.zeros(()))
        p = torch.softmax(s / D ** 0.5, dim=-1)
        o = p @ v[b, :, :L]
        pw = torch.zeros(n_heads, L, 2 * window + 1)
        for r in range(-window, window + 1):
            ii = torch.arange(max(0, -r), min(L, L - r))
            if len(ii):
                pw[:, ii, r + window] = p[:, ii, ii + r]
        out[b, :, :L] = o + pw @ ev
    return out.transpose(1, 2).reshape(B, T, HD).to(out_dtype)


def conformer_attention(q, k, v, pos, u_bias, v_bias, n_heads, lens, out_dtype):
    B, T, HD = q.shape
    D = HD // n_heads
    out = torch.zeros(B, T, HD)
    for b in range(B):
        L = T if lens is None else int(lens[b])
        if L == 0:
            continue
        qq = q[b, :L].float().reshape(L, n_heads, D)
        kk = k[b, :L].float().reshape(L, n_heads, D).permute(1, 0, 2)
        vv = v[b, :L].float().reshape(L, n_heads, D).permute(1, 0, 2)
        pp = pos[:L].float().reshape(L, n_heads, D)
        content = (qq + u_bias.float()).transpose(0, 1) @ kk.transpose(1, 2)          # [H, L, L]
        ps = (qq + v_bias.float()).transpose(0, 1) @ pp.permute(1, 2, 0)               # [H, L, L]
        padded = torch.cat([ps.new_zeros(n_heads, L, 1), ps], dim=-1).reshape(n_heads, L + 1, L)
        ps = padded[:, 1:].reshape(n_heads, L, L)
        a = torch.softmax((content + ps) / HD ** 0.5, dim=-1)
        out[b, :L] = (a @ vv).transpose(0, 1).reshape(L, HD)
    return out.to(out_dtype)


def instnorm_stats(x, lens, eps=1e-5):
    B, T, C = x.shape
    st = torch.zeros(B, C, 2)
    for b in range(B):
        L = T if lens is None else int(lens[b])
        xb = x[b, :L].float()
        st[b, :, 0] = xb.mean(0)
        st[b, :, 1] = torch.rsqrt(xb.var(0, unbiased=False) + eps)
    return st


def adain_apply(x, stats, gb, slope, lens, out_dtype, up_w=None, up_b=None, out=None):
    B, T, C = x.shape
    g, be = gb[:, :C].float(), gb[:, C:].float()
    a = (x.float() - stats[:, None, :, 0]) * stats[:, None, :, 1] * (1 + g[:, None]) + be[:, None]
    a = torch.where(a > 0, a, a * slope)
    a = _mask_rows(a, lens, T)
    if up_w is not None:
        w = up_w.float().reshape(C, 3)
        an = torch.cat([a[:, 1:], a.new_zeros(B, 1, C)], dim=1)
        even = a * w[:, 1] + up_b.float()
        odd = a * w[:, 2] + an * w[:, 0] + up_b.float()
        a = torch.stack([even, odd], dim=2).reshape(B, 2 * T, C)
        a = _mask_rows(a, None if lens is None else lens * 2, 2 * T)
    if out is None:
        return a.to(out_dtype)
    out.copy_(a.to(out.dtype))
    return out


def adain_norm(x, gb, slope, lens, out_dtype, up_w=None, up_b=None, eps=1e-5):
    return adain_apply(x, instnorm_stats(x, lens, eps), gb, slope, lens, out_dtype, up_w, up_b)


def repeat_rows(x, rep, lens, out_dtype=None, out=None):
    y = _mask_rows(x.float(), lens, x.shape[1]).repeat_interleave(rep, dim=1)
    if out is None:
        return y.to(out_dtype or x.dtype)
    out.copy_(y.to(out.dtype))
    return out


def length_regulate(x, dur, lens_t, rep, To, out=None, out_dtype=None):
    B, Tt, C = x.shape
    y = torch.zeros(B, To, C)
    olens = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        nt = Tt if lens_t is None else int(lens_t[b])
        d = dur[b, :nt].long().clamp(min=0)
        idx = torch.repeat_interleave(torch.arange(nt), d * rep)[:To]
        y[b, :len(idx)] = x[b].float()[idx]
        olens[b] = min(int(d.sum()) * rep, To)
    if out is None:
        return y.to(out_dtype or x.dtype), olens
    out.copy_(y.to(out.dtype))
    return out, olens


def round_durations(pred, lens_t):
    B, Tt = pred.shape
    d = torch.round(pred.float()).clamp(min=1).to(torch.int32)
    if lens_t is not None:
        d = d * (torch.arange(Tt)[None, :] < lens_t[:, None].long()).to(torch.int32)
    return d, d.sum(dim=1).to(torch.int32)


def conv_small(x, sc, *, raw=None, act_out=None, act=ops.ACT_NONE, slope=0.0, lens=None):
    squeeze = x.dim() == 3
    x4 = (x.unsqueeze(2) if squeeze else x).float()
    B, T, Fd, Cin = x4.shape
    w = sc.w.float()
    acc = torch.zeros(B, T, Fd, w.shape[1])
    for j, (dt, df) in enumerate(sc.taps):
        t_idx, f_idx = torch.arange(T) + dt, torch.arange(Fd) + df
        tv, fv = (t_idx >= 0) & (t_idx < T), (f_idx >= 0) & (f_idx < Fd)
        g = x4[:, t_idx.clamp(0, T - 1)][:, :, f_idx.clamp(0, Fd - 1)]
        g = g * (tv[None, :, None, None] & fv[None, None, :, None])
        acc += g @ w[j].t()
    if sc.bias is not None:
        acc = acc + sc.bias.float()
    acc = _mask_rows(acc, lens, T)
    if squeeze:
        acc = acc.squeeze(2)
    return _write(raw, acc.shape, acc), _write(act_out, acc.shape, _ACT[act](acc, slope))


def dwconv(x, w, bias, k, stride, pad, *, glu=False, act=ops.ACT_NONE, slope=0.0, out_dtype=None,
           lens_in=None, lens_out=None):
    squeeze = x.dim() == 3
    x4 = (x.unsqueeze(2) if squeeze else x).float()
    if glu:
        C = x4.shape[-1] // 2
        x4 = x4[..., :C] * torch.sigmoid(x4[..., C:])
    B, T, Fd, C = x4.shape
    x4 = _mask_rows(x4, lens_in, T)
    (kt, kf) = k
    wt = w.float().reshape(kt, kf, C).permute(2, 0, 1).unsqueeze(1)     # [C,1,kt,kf]
    y = F.conv2d(x4.permute(0, 3, 1, 2), wt, None if bias is None else bias.float(), stride=stride,
                 padding=pad, groups=C).permute(0, 2, 3, 1)
    y = _ACT[act](y, slope)
    y = _mask_rows(y, lens_out, y.shape[1])
    if squeeze:
        y = y.squeeze(2)
    return y.to(out_dtype or x.dtype)


def avgpool(x, pt, pf, out_dtype=None):
    squeeze = x.dim() == 3
    x4 = (x.unsqueeze(2) if squeeze else x).float()
    B, T, Fd, C = x4.shape
    if T % pt:
        x4 = torch.cat([x4] + [x4[:, -1:]] * (pt - T % pt), dim=1)
    y = F.avg_pool2d(x4.permute(0, 3, 1, 2), (pt, pf)).permute(0, 2, 3, 1)
    if squeeze:
        y = y.squeeze(2)
    return y.to(out_dtype or x.dtype)


def affine_act_maxpool(x, scale, shift, slope, pf, out_dtype):
    y = x.float() * scale.float() + shift.float()
    y = torch.where(y > 0, y, y * slope)
    y = F.max_pool2d(y.permute(0, 3, 1, 2), (1, pf)).permute(0, 2, 3, 1)
    return y.to(out_dtype)


def global_avgpool(x, slope, out_dtype, t_stride=1):
    x4 = (x.unsqueeze(2) if x.dim() == 3 else x).float()[:, ::t_stride]
    y = torch.where(x4 > 0, x4, x4 * slope).mean(dim=(1, 2))
    return y.to(out_dtype)


# tests that check the host logic to fp32 rounding noise switch this off
EMULATE_FP16_RECURRENCE = True


def bilstm(xproj, whh_t, hidden, lens, out_dtype):
    B, T, _ = xproj.shape
    H = hidden
    out = torch.zeros(B, T, 2 * H)
    for b in range(B):
        L = T if lens is None else int(lens[b])
        for d in range(2):
            h = torch.zeros(H)
            c = torch.zeros(H)
            W = whh_t[d].float()                                      # [H, 4H]
            q16 = H in (128, 256) and EMULATE_FP16_RECURRENCE
            if q16:   # the H=128/256 kernels keep W_hh and h as fp16 tensor-core operands (fp32 accumulate)
                W = W.half().float()
            order = range(L) if d == 0 else range(L - 1, -1, -1)
            for t in order:
                hq = h.half().float() if q16 else h
                g = xproj[b, t, d * 4 * H:(d + 1) * 4 * H].float() + hq @ W
                i, f, gg, o = g.split(H)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                out[b, t, d * H:(d + 1) * H] = h
    return out.to(out_dtype)


def lstm_onestep(xproj, hidden, out_dtype):
    H = hidden
    outs = []
    for d in range(2):
        g = xproj[..., d * 4 * H:(d + 1) * 4 * H].float()
        i, f, gg, o = g.split(H, dim=-1)
        c = torch.sigmoid(i) * torch.tanh(gg)
        outs.append(torch.sigmoid(o) * torch.tanh(c))
    return torch.cat(outs, dim=-1).to(out_dtype)


def log_norm(mel):
    return torch.log(torch.exp(mel.float() * 4 - 4).norm(dim=1))


def to_channels_last(src, out_dtype, lens=None, sub=None, mul=None, out=None):
    y = src.float()
    if sub is not None:
        y = (y - sub.float()[None, :, None]) * mul.float()[None, :, None]
    y = _mask_rows(y.transpose(1, 2), lens, src.shape[2])
    if out is None:
        return y.to(out_dtype).contiguous()
    out.copy_(y.to(out.dtype))
    return out


def to_channels_first(src, out_dtype, lens=None, sub=None, mul=None):
    y = src.float()
    if sub is not None:
        y = (y - sub.float()) * mul.float()
    y = _mask_rows(y, lens, src.shape[1])
    return y.transpose(1, 2).to(out_dtype).contiguous()


SIM_FUNCS = ["conv", "resblock_pair", "embed", "layernorm", "relpos_attention", "conformer_attention", "instnorm_stats",
             "adain_apply", "adain_norm", "repeat_rows", "length_regulate", "round_durations", "conv_small", "dwconv", "avgpool",
             "affine_act_maxpool", "global_avgpool", "bilstm", "lstm_onestep", "log_norm",
             "to_channels_last", "to_channels_first"]


def install(monkeypatch):
    g = globals()
    for name in SIM_FUNCS:
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(ops, "_require_cuda", lambda t, name: None)
