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


# --------------------------------------------------------------------------------------------
# Utils/RelTransformerEnc.py
# --------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    """View of ``sd`` restricted to ``prefix`` (keys with the prefix stripped)."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def channel_ln(x, gamma, beta, eps=1e-4):
    """LayerNorm over dim 1 of [B,C,T] (RelTransformerEnc.py:272-290)."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def relpos_attention(sd, pre, x, mask, n_heads=4, window=4):
    """MultiHeadAttention.forward/attention (RelTransformerEnc.py:127-169) in the closed form of
    SURVEY.md Appendix A (verified equal to the pad/view skew implementation)."""
    B, C, T = x.shape
    D = C // n_heads
    proj = lambda name, t: F.conv1d(t, sd[f"{pre}.{name}.weight"].float(), sd[f"{pre}.{name}.bias"].float())
    q, k, v = (proj(n_, x).view(B, n_heads, D, T).transpose(2, 3) for n_ in ("conv_q", "conv_k", "conv_v"))
    ek, ev = sd[f"{pre}.emb_rel_k"].float()[0], sd[f"{pre}.emb_rel_v"].float()[0]      # [2w+1, D]
    scores = q @ k.transpose(-2, -1)
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]                                                   # j - i
    inwin = rel.abs() <= window
    qe = q @ ek.t()                                                                     # [B,H,T,2w+1]
    gathered = torch.gather(qe, 3, (rel + window).clamp(0, 2 * window).expand(B, n_heads, T, T))
    scores = (scores + gathered * inwin) / math.sqrt(D)
    scores = scores.masked_fill(mask == 0, -1e4)
    p = torch.softmax(scores, dim=-1)
    out = p @ v
    pw = torch.zeros(B, n_heads, T, 2 * window + 1)
    for r in range(-window, window + 1):
        ii = torch.arange(max(0, -r), min(T, T - r))
        if len(ii):
            pw[:, :, ii, r + window] = p[:, :, ii, ii + r]
    out = out + pw @ ev
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return proj("conv_o", out)


@torch.no_grad()
def rel_transformer_encoder(sd, tokens, lengths, n_layers, hidden=512, n_heads=4, window=4, kernel=9):
    """RelTransformerEncoder.forward (RelTransformerEnc.py:371-380) -> [B,T,H]."""
    T = tokens.shape[1]
    x = F.embedding(tokens, sd["emb.weight"].float()) * math.sqrt(hidden)
    x = x.transpose(1, 2)
    m = (torch.arange(T)[None, :] < lengths[:, None]).unsqueeze(1).float()                  # [B,1,T]
    # prenet ConvReluNorm (:316-325)
    x0 = x
    for i in range(3):
        x = F.conv1d(x * m, sd[f"pre.conv_layers.{i}.weight"].float(), sd[f"pre.conv_layers.{i}.bias"].float(), padding=2)
        x = torch.relu(channel_ln(x, sd[f"pre.norm_layers.{i}.gamma"].float(), sd[f"pre.norm_layers.{i}.beta"].float()))
    x = (x0 + F.conv1d(x, sd["pre.proj.weight"].float(), sd["pre.proj.bias"].float())) * m
    # Encoder, pre-LN (:67-90)
    am = m.unsqueeze(2) * m.unsqueeze(-1)
    for i in range(n_layers):
        x = x * m
        y = channel_ln(x, sd[f"encoder.norm_layers_1.{i}.gamma"].float(), sd[f"encoder.norm_layers_1.{i}.beta"].float())
        x = x + relpos_attention(sd, f"encoder.attn_layers.{i}", y, am, n_heads, window)
        y = channel_ln(x, sd[f"encoder.norm_layers_2.{i}.gamma"].float(), sd[f"encoder.norm_layers_2.{i}.beta"].float())
        y = F.conv1d(y * m, sd[f"encoder.ffn_layers.{i}.conv_1.weight"].float(),
                     sd[f"encoder.ffn_layers.{i}.conv_1.bias"].float(), padding=kernel // 2)
        y = F.conv1d(torch.relu(y) * m, sd[f"encoder.ffn_layers.{i}.conv_2.weight"].float(),
                     sd[f"encoder.ffn_layers.{i}.conv_2.bias"].float())
        x = x + y * m
    x = channel_ln(x, sd["encoder.last_ln.gamma"].float(), sd["encoder.last_ln.beta"].float()) * m
    return x.transpose(1, 2)


# --------------------------------------------------------------------------------------------
# models.py building blocks
# --------------------------------------------------------------------------------------------
SQ2 = math.sqrt(2.0)


def adain(sd, pre, x, s):
    """AdaIN1d (models.py:230-240): (1+gamma) * InstanceNorm(x) + beta."""
    h = F.linear(s, sd[f"{pre}.fc.weight"].float(), sd[f"{pre}.fc.bias"].float())
    gamma, beta = h.chunk(2, dim=1)
    return (1 + gamma.unsqueeze(-1)) * F.instance_norm(x, eps=1e-5) + beta.unsqueeze(-1)


def adain_resblk1d(sd, pre, x, s, upsample=False):
    """AdainResBlk1d.forward (models.py:183-202)."""
    r = F.leaky_relu(adain(sd, f"{pre}.norm1", x, s), 0.2)
    if upsample:
        C = x.shape[1]
        r = F.conv_transpose1d(r, wn_weight(sd, f"{pre}.pool"), sd[f"{pre}.pool.bias"].float(), stride=2, padding=1,
                               output_padding=1, groups=C)
    r = F.conv1d(r, wn_weight(sd, f"{pre}.conv1"), sd[f"{pre}.conv1.bias"].float(), padding=1)
    r = F.leaky_relu(adain(sd, f"{pre}.norm2", r, s), 0.2)
    r = F.conv1d(r, wn_weight(sd, f"{pre}.conv2"), sd[f"{pre}.conv2.bias"].float(), padding=1)
    sc = F.interpolate(x, scale_factor=2, mode="nearest") if upsample else x
    if f"{pre}.conv1x1.weight_v" in sd or f"{pre}.conv1x1.weight" in sd:
        sc = F.conv1d(sc, wn_weight(sd, f"{pre}.conv1x1"))
    return (r + sc) / SQ2


def _avgpool_rep(x, k):
    """DownSample (models.py:43-57): replicate the last column when the last dim is odd."""
    if x.shape[-1] % 2 != 0:
        x = torch.cat([x, x[..., -1:]], dim=-1)
    return F.avg_pool2d(x, k) if x.dim() == 4 else F.avg_pool1d(x, k)


def resblk2d(sd, pre, x, down):
    """ResBlk.forward (models.py:79-100), normalize=False."""
    k, s, p = {"half": ((3, 3), (2, 2), (1, 1)), "channelpreserve": ((1, 3), (1, 2), (0, 1))}[down]
    C = x.shape[1]
    r = F.conv2d(F.leaky_relu(x, 0.2), sn_weight(sd, f"{pre}.conv1"), sd[f"{pre}.conv1.bias"].float(), padding=1)
    r = F.conv2d(r, sn_weight(sd, f"{pre}.downsample_res.conv"), sd[f"{pre}.downsample_res.conv.bias"].float(),
                 stride=s, padding=p, groups=C)
    r = F.conv2d(F.leaky_relu(r, 0.2), sn_weight(sd, f"{pre}.conv2"), sd[f"{pre}.conv2.bias"].float(), padding=1)
    sc = x
    if f"{pre}.conv1x1.weight_orig" in sd:
        sc = F.conv2d(sc, sn_weight(sd, f"{pre}.conv1x1"))
    sc = _avgpool_rep(sc, 2 if down == "half" else (1, 2))
    return (sc + r) / SQ2


def resblk1d(sd, pre, x):
    """ResBlk1d.forward with downsample=True (models.py:127-156)."""
    C = x.shape[1]
    r = F.conv1d(F.leaky_relu(x, 0.2), wn_weight(sd, f"{pre}.conv1"), sd[f"{pre}.conv1.bias"].float(), padding=1)
    r = F.conv1d(r, wn_weight(sd, f"{pre}.pool"), sd[f"{pre}.pool.bias"].float(), stride=2, padding=1, groups=C)
    r = F.conv1d(F.leaky_relu(r, 0.2), wn_weight(sd, f"{pre}.conv2"), sd[f"{pre}.conv2.bias"].float(), padding=1)
    sc = x
    if f"{pre}.conv1x1.weight_v" in sd:
        sc = F.conv1d(sc, wn_weight(sd, f"{pre}.conv1x1"))
    return (_avgpool_rep(sc, 2) + r) / SQ2


def style_stack_2d(sd, pre, img, downs, last_idx, last_stride):
    """Mel_block / EMA_block / dur_block Sequential (models.py:385-401,530-537) -> [B,C]."""
    x = F.conv2d(img, sn_weight(sd, f"{pre}.0"), sd[f"{pre}.0.bias"].float(), padding=1)
    for i, d in enumerate(downs):
        x = resblk2d(sd, f"{pre}.{i + 1}", x, d)
    x = F.conv2d(F.leaky_relu(x, 0.2), sn_weight(sd, f"{pre}.{last_idx}"), sd[f"{pre}.{last_idx}.bias"].float(),
                 stride=last_stride)
    return F.leaky_relu(x, 0.2).mean(dim=(2, 3))


def style_stack_1d(sd, pre, x):
    """F0_block / energy_block (models.py:402-411) -> [B,C]."""
    x = F.conv1d(x, sn_weight(sd, f"{pre}.0"), sd[f"{pre}.0.bias"].float(), padding=1)
    for i in range(1, 5):
        x = resblk1d(sd, f"{pre}.{i}", x)
    return F.leaky_relu(x, 0.2).mean(dim=2)


def bn_eval(sd, pre, x):
    shape = [1, -1] + [1] * (x.dim() - 2)
    scale = sd[f"{pre}.weight"].float() / torch.sqrt(sd[f"{pre}.running_var"].float() + 1e-5)
    return (x - sd[f"{pre}.running_mean"].float().view(shape)) * scale.view(shape) + sd[f"{pre}.bias"].float().view(shape)


def bilstm(sd, pre, x, lengths=None):
    """Single-layer bidirectional LSTM, batch_first (PyTorch gate order i,f,g,o) run through
    torch's own nn.LSTM with the state-dict's weights, like the reference does; with ``lengths`` it
    is wrapped in pack_padded_sequence / pad_packed_sequence (models.py:555-564)."""
    H = sd[f"{pre}.weight_hh_l0"].shape[1]
    lstm = torch.nn.LSTM(x.shape[-1], H, 1, batch_first=True, bidirectional=True)
    lstm.load_state_dict({k: sd[f"{pre}.{k}"].float() for k in lstm.state_dict()})
    if lengths is None:
        return lstm(x)[0]
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), batch_first=True, enforce_sorted=False)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=x.shape[1])
    return out


# --------------------------------------------------------------------------------------------
# Utils/JDC/model.py
# --------------------------------------------------------------------------------------------
def jdc_resblock(sd, pre, x):
    """ResBlock.forward (Utils/JDC/model.py:183-190)."""
    x = F.max_pool2d(F.leaky_relu(bn_eval(sd, f"{pre}.pre_conv.0", x), 0.01), (1, 2))
    y = F.conv2d(x, sd[f"{pre}.conv.0.weight"].float(), padding=1)
    y = F.leaky_relu(bn_eval(sd, f"{pre}.conv.1", y), 0.01)
    y = F.conv2d(y, sd[f"{pre}.conv.3.weight"].float(), padding=1)
    return y + F.conv2d(x, sd[f"{pre}.conv1by1.weight"].float())


@torch.no_grad()
def jdc_forward(sd, mel4):
    """JDCNet.forward (Utils/JDC/model.py:102-137): [B,1,80,T] -> [B,1,T]."""
    T = mel4.shape[-1]
    x = mel4.float().transpose(-1, -2)
    x = F.conv2d(x, sd["conv_block.0.weight"].float(), padding=1)
    x = F.leaky_relu(bn_eval(sd, "conv_block.1", x), 0.01)
    x = F.conv2d(x, sd["conv_block.3.weight"].float(), padding=1)
    for i in (1, 2, 3):
        x = jdc_resblock(sd, f"res_block{i}", x)
    x = F.max_pool2d(F.leaky_relu(bn_eval(sd, "pool_block.0", x), 0.01), (1, 4))
    x = x.permute(0, 2, 1, 3).contiguous().view(-1, T, 512)
    x = bilstm(sd, "bilstm_classifier", x)
    x = F.linear(x, sd["classifier.weight"].float(), sd["classifier.bias"].float())
    return x.abs().transpose(-1, -2)


# --------------------------------------------------------------------------------------------
# Utils/EMA/EMA_Predictor.py + conformer
# --------------------------------------------------------------------------------------------
def _lin(sd, pre, x):
    return F.linear(x, sd[f"{pre}.weight"].float(), sd[f"{pre}.bias"].float() if f"{pre}.bias" in sd else None)


def conformer_block(sd, pre, x):
    """ConformerBlock (conformer/encoder.py:74-110) on [B,T,256]."""
    D, H = 256, 4

    def ff(p, t):                                                              # feed_forward.py:47-57
        t = F.layer_norm(t, (D,), sd[f"{p}.0.weight"].float(), sd[f"{p}.0.bias"].float())
        t = _lin(sd, f"{p}.1.linear", t)
        return _lin(sd, f"{p}.4.linear", t * torch.sigmoid(t))

    x = x + 0.5 * ff(f"{pre}.sequential.0.module.sequential", x)
    # relative MHSA (attention.py:72-113)
    mp = f"{pre}.sequential.1.module"
    B, T, _ = x.shape
    y = F.layer_norm(x, (D,), sd[f"{mp}.layer_norm.weight"].float(), sd[f"{mp}.layer_norm.bias"].float())
    pe = sd[f"{mp}.positional_encoding.pe"].float()[:, :T].repeat(B, 1, 1)
    ap = f"{mp}.attention"
    q = _lin(sd, f"{ap}.query_proj.linear", y).view(B, T, H, D // H)
    k = _lin(sd, f"{ap}.key_proj.linear", y).view(B, T, H, D // H).permute(0, 2, 1, 3)
    v = _lin(sd, f"{ap}.value_proj.linear", y).view(B, T, H, D // H).permute(0, 2, 1, 3)
    p = F.linear(pe, sd[f"{ap}.pos_proj.linear.weight"].float()).view(B, T, H, D // H)
    content = (q + sd[f"{ap}.u_bias"].float()).transpose(1, 2) @ k.transpose(2, 3)
    pos = (q + sd[f"{ap}.v_bias"].float()).transpose(1, 2) @ p.permute(0, 2, 3, 1)
    padded = torch.cat([pos.new_zeros(B, H, T, 1), pos], dim=-1).view(B, H, T + 1, T)   # _relative_shift
    pos = padded[:, :, 1:].view_as(pos)
    attn = torch.softmax((content + pos) / math.sqrt(D), dim=-1)
    ctx = (attn @ v).transpose(1, 2).contiguous().view(B, T, D)
    x = x + _lin(sd, f"{ap}.out_proj.linear", ctx)
    # convolution module (convolution.py:136-149)
    cp = f"{pre}.sequential.2.module.sequential"
    y = F.layer_norm(x, (D,), sd[f"{cp}.0.weight"].float(), sd[f"{cp}.0.bias"].float()).transpose(1, 2)
    y = F.conv1d(y, sd[f"{cp}.2.conv.weight"].float(), sd[f"{cp}.2.conv.bias"].float())
    a, g = y.chunk(2, dim=1)
    y = a * torch.sigmoid(g)
    y = F.conv1d(y, sd[f"{cp}.4.conv.weight"].float(), None, padding=15, groups=D)
    y = bn_eval(sd, f"{cp}.5", y)
    y = y * torch.sigmoid(y)
    y = F.conv1d(y, sd[f"{cp}.7.conv.weight"].float(), sd[f"{cp}.7.conv.bias"].float()).transpose(1, 2)
    x = x + y
    x = x + 0.5 * ff(f"{pre}.sequential.3.module.sequential", x)
    return F.layer_norm(x, (D,), sd[f"{pre}.sequential.4.weight"].float(), sd[f"{pre}.sequential.4.bias"].float())


@torch.no_grad()
def ema_predictor(sd, f0, energy, mel):
    """EMA_Predictor.forward (Utils/EMA/EMA_Predictor.py:65-82) at batch 1 semantics per item:
    decoder2 is an nn.LSTM without batch_first fed [B,T,256], i.e. at B=1 one step from zero state
    per frame — restated as that single step."""
    x = torch.cat((f0, energy, mel), 1).transpose(1, 2)
    x = _lin(sd, "encoder1.0", x)
    x = torch.relu(bn_eval(sd, "encoder1.2", x.transpose(1, 2)).transpose(1, 2))
    for i in range(3):
        x = conformer_block(sd, f"decoder.{i}", x)
    outs = []
    for sfx in ("", "_reverse"):
        g = F.linear(x, sd[f"decoder2.weight_ih_l0{sfx}"].float(),
                     sd[f"decoder2.bias_ih_l0{sfx}"].float() + sd[f"decoder2.bias_hh_l0{sfx}"].float())
        i, f, gg, o = g.chunk(4, dim=-1)
        c = torch.sigmoid(i) * torch.tanh(gg)
        outs.append(torch.sigmoid(o) * torch.tanh(c))
    x = torch.cat(outs, dim=-1)
    x = _lin(sd, "decoder3.0", x)
    x = torch.relu(bn_eval(sd, "decoder3.2", x.transpose(1, 2)).transpose(1, 2))
    return _lin(sd, "decoder3.5", x).transpose(1, 2)


# --------------------------------------------------------------------------------------------
# models.py modules
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def style_encoder(sd, mel, dist):
    """StyleEncoder.forward, eval branch, batch 1 (models.py:426-433,447-472).
    ``mel`` [1,80,T] -> (f0 [1,1,T], n [1,1,T], ema [1,10,T], Style [1,512])."""
    n = torch.log(torch.exp(mel.unsqueeze(1) * 4 - 4).norm(dim=2))                     # log_norm (:655-660)
    f0 = jdc_forward(_sub(sd, "pitch_extractor."), mel.unsqueeze(1))
    ema = ema_predictor(_sub(sd, "ema_extractor."), f0, n, mel)
    n = (n - dist["energy_mean"]) / dist["energy_std"]
    f0 = (f0 - dist["pitch_mean"]) / dist["pitch_std"]
    ema = ((ema.transpose(1, 2) - dist["EMA_mean"]) / dist["EMA_std"]).transpose(1, 2)
    Tc = mel.shape[2] - 1                                                              # crop, start 0 (:459-466)
    heads = [
        (style_stack_2d(sd, "Mel_block", mel[:, None, :, :Tc], ["half"] * 4, 6, 1), "Mellinear"),
        (style_stack_2d(sd, "EMA_block", ema[:, None, :, :Tc], ["channelpreserve", "channelpreserve", "half"], 5, 2), "EMAlinear"),
        (style_stack_1d(sd, "F0_block", f0[:, :, :Tc]), "F0linear"),
        (style_stack_1d(sd, "energy_block", n[:, :, :Tc]), "Energylinear")]
    style = torch.cat([_lin(sd, name, feat) for feat, name in heads], dim=1)
    return f0, n, ema, style


@torch.no_grad()
def duration_predictor(sd, tokens, ema, lengths):
    """DurationPredictor.forward (models.py:540-566), batch 1 -> [1,Tt]."""
    dstyle = _lin(sd, "dur_linear", style_stack_2d(sd, "dur_block", ema[:, None], ["channelpreserve", "channelpreserve", "half"], 5, 2))
    d = rel_transformer_encoder(_sub(sd, "text_encoder."), tokens, lengths, n_layers=2).transpose(1, 2)
    for i in range(3):
        d = adain_resblk1d(sd, f"duration.{i}", d, dstyle)
    x = bilstm(sd, "LSTM", d.transpose(1, 2), lengths)
    return _lin(sd, "duration_proj.linear_layer", x).squeeze(-1)


@torch.no_grad()
def arts_predictor(sd, a_en, style):
    """ArtsPredictor.forward (models.py:596-621)."""
    sl = {"F0": style[:, 384:448], "N": style[:, 448:512], "EMA": style[:, 256:384]}
    x = adain_resblk1d(sd, "shared", a_en, style)
    outs = []
    for name in ("F0", "N", "EMA"):
        y = adain_resblk1d(sd, f"{name}.0", x, style, upsample=True)
        for j in (1, 2):
            y = adain_resblk1d(sd, f"{name}.{j}", y, sl[name])
        y = bilstm(sd, f"{name}_LSTM", y.transpose(1, 2)).transpose(1, 2)
        outs.append(F.conv1d(y, sd[f"{name}_proj.weight"].float(), sd[f"{name}_proj.bias"].float()))
    return outs


@torch.no_grad()
def decoder(sd, asr, style, f0, n, ema):
    """Decoder.forward (models.py:497-517)."""
    asr = F.interpolate(asr, scale_factor=2, mode="nearest")
    c1 = lambda name, t: F.conv1d(t, wn_weight(sd, name), sd[f"{name}.bias"].float())
    f0, n, ema = c1("F0_conv", f0), c1("N_conv", n), c1("EMA_conv", ema)
    x = adain_resblk1d(sd, "encode", torch.cat([asr, f0, n, ema], 1), style)
    res = c1("asr_res.0", asr)
    for i in range(3):
        x = adain_resblk1d(sd, f"decode.{i}", torch.cat([x, res, f0, n, ema], 1), style)
    for i in range(3, 6):
        x = adain_resblk1d(sd, f"decode.{i}", x, style[:, :256])
    return c1("to_out.0", x)


@torch.no_grad()
def artsspeech_test(sd, tokens, mel, dist, durations=None, want_aux=False, predict_durations=False):
    """ArtsSpeech.forward(step='test') (models.py:356-371) for ONE utterance:
    ``tokens`` [1,Tt], ``mel`` [1,80,Tr] -> mel [1,80,2*sum(dur)].  ``durations`` (int [Tt]) replaces
    round(duration).clamp(min=1); with ``predict_durations`` the duration predictor (:360) still runs (its
    output is returned in the aux dict) while the given durations drive the length regulation -- the
    benchmark's step, identical in both arms."""
    lengths = torch.tensor([tokens.shape[1]])
    t_en = rel_transformer_encoder(_sub(sd, "text_encoder."), tokens, lengths, 4).transpose(1, 2)
    a_en = rel_transformer_encoder(_sub(sd, "arts_encoder."), tokens, lengths, 4).transpose(1, 2)
    f0e, ne, emae, style = style_encoder(_sub(sd, "style_encoder."), mel, dist)
    duration = None
    if durations is None or predict_durations:
        duration = duration_predictor(_sub(sd, "durationPredictor."), tokens, emae, lengths)
    if durations is None:
        durations = torch.round(duration.squeeze(0)).clamp(min=1)
    idx = torch.repeat_interleave(torch.arange(tokens.shape[1]), durations.long().view(-1))   # one-hot matmul == gather
    f0, n, ema = arts_predictor(_sub(sd, "artsPredictor."), a_en[:, :, idx], style)
    out = decoder(_sub(sd, "decoder."), t_en[:, :, idx], style, f0, n, ema)
    if want_aux:
        return out, dict(T_en=t_en.transpose(1, 2), style=style, duration=duration, pred_dur=durations.long(),
                         F0=f0, N=n, EMA=ema, f0_ext=f0e, n_ext=ne, ema_ext=emae)
    return out
