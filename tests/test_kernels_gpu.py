"""GPU parity of every C-ABI primitive against its torch semantic model (tests/sim_backend.py).

The semantic models are what the CPU host-logic tests run the module classes on, so agreement here
closes the loop: module(sim) == oracle on CPU, kernel == sim on GPU.
"""
import pytest
import torch

from artspeech_b200 import ops
from tests import sim_backend as sim

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, tol, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    scale = max(b.abs().max().item(), 1.0)
    assert err <= tol * scale, f"{what}: max err {err:.3e} (scale {scale:.2f})"


def _lens(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    l = torch.randint(max(T // 3, 1), T + 1, (B,), generator=g, dtype=torch.int32)
    l[0] = T
    return l


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 150, 512, 512, 5, 1), (2, 300, 128, 128, 7, 3), (1, 1000, 32, 32, 11, 5),
                                   (3, 77, 640, 1024, 3, 1), (2, 64, 80, 512, 7, 1), (2, 40, 88, 256, 1, 1),
                                   (1, 500, 32, 1, 7, 1), (2, 130, 64, 192, 3, 1),
                                   (2, 1000, 64, 64, 7, 3), (2, 500, 128, 128, 3, 1), (1, 700, 64, 64, 11, 5),
                                   (2, 300, 64, 32, 3, 1), (3, 257, 128, 64, 3, 1)])
def test_conv_igemm_1d(shape, dt):
    B, T, Cin, Cout, k, dil = shape
    torch.manual_seed(1)
    x = (torch.randn(B, T, Cin) * 0.5).to(dt)
    w = torch.randn(k, Cout, Cin) / (Cin * k) ** 0.5
    bias = torch.randn(Cout)
    r1 = torch.randn(B, T, Cout).to(dt)
    r2 = torch.randn(B, T, Cout)
    lens = _lens(B, T)
    pw_c = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, "cpu")
    pw_g = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, DEV)
    ref_raw, ref_act = sim.conv(x, pw_c, res1=r1, res2=r2, scale=0.7, raw=torch.float32, act_out=dt,
                                act=ops.ACT_LRELU, slope=0.1, lens=lens)
    raw, act = ops.conv(x.to(DEV), pw_g, res1=r1.to(DEV), res2=r2.to(DEV), scale=0.7, raw=torch.float32,
                        act_out=dt, act=ops.ACT_LRELU, slope=0.1, lens=lens.to(DEV))
    _close(raw, ref_raw, 2e-4, "raw")
    _close(act, ref_act, 1e-2 if dt == torch.bfloat16 else 2e-3, "act")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 1000, 32, 7, 1), (1, 700, 64, 5, 3), (3, 257, 16, 3, 1), (1, 255, 32, 1, 1)])
def test_conv_single_output_channel(shape, dt):
    """Cout = 1 convolutions (the vocoder's conv_post + tanh) take the streaming kernel in conv_cout1.cu."""
    B, T, Cin, k, dil = shape
    torch.manual_seed(4)
    x = (torch.randn(B, T, Cin) * 0.5).to(dt)
    w = torch.randn(k, 1, Cin) / (Cin * k) ** 0.5
    bias = torch.randn(1)
    lens = _lens(B, T)
    pw_c = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, "cpu")
    pw_g = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, DEV)
    ref_raw, ref_act = sim.conv(x, pw_c, scale=0.9, raw=torch.float32, act_out=torch.float32, act=ops.ACT_TANH, lens=lens)
    raw, act = ops.conv(x.to(DEV), pw_g, scale=0.9, raw=torch.float32, act_out=torch.float32, act=ops.ACT_TANH,
                        lens=lens.to(DEV))
    _close(raw, ref_raw, 2e-4, "raw")
    _close(act, ref_act, 2e-4, "act")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 700, 64, 3, 1), (2, 1000, 64, 7, 3), (1, 1500, 64, 11, 5), (3, 333, 32, 3, 5),
                                   (2, 2000, 32, 11, 5), (1, 900, 32, 7, 1), (2, 600, 128, 3, 3), (1, 800, 128, 7, 5),
                                   (1, 700, 128, 11, 5), (4, 100, 64, 7, 5), (1, 246, 32, 11, 1), (1, 247, 64, 11, 3)])
@pytest.mark.parametrize("variant", ["mid", "branch_end", "stage_end"])
def test_resblock_pair(shape, variant, dt):
    """as_hifigan_resblock_pair vs its contract model (Vocoder/vocoder.py:35-42 with the activated stream)."""
    B, L, C, k, dil = shape
    torch.manual_seed(2)
    x_raw = torch.randn(B, L, C)
    lens = _lens(B, L)
    x_raw = x_raw * (torch.arange(L)[None, :] < lens[:, None])[:, :, None]
    xa = torch.where(x_raw > 0, x_raw, 0.1 * x_raw).to(dt)
    w1 = torch.randn(k, C, C) / (C * k) ** 0.5
    w2 = torch.randn(k, C, C) / (C * k) ** 0.5
    b1, b2 = torch.randn(C) * 0.1, torch.randn(C) * 0.1
    kw = dict(slope=0.1)
    if variant == "mid":
        kw.update(out_act=ops.ACT_LRELU, out_slope=0.1)
    elif variant == "stage_end":
        kw.update(res2=torch.randn(B, L, C).to(dt), res3=torch.randn(B, L, C).to(dt), scale=1.0 / 3,
                  out_act=ops.ACT_LRELU, out_slope=0.01)
    packs = {}
    for dev in ("cpu", DEV):
        packs[dev] = (ops.pack_conv(w1, b1, ops.taps_1d(k, dil), dt, dev), ops.pack_conv(w2, b2, ops.taps_1d(k, 1), dt, dev))
    ref = sim.resblock_pair(xa, *packs["cpu"], k, dil, lens=lens, **kw)
    kw_g = {n: (v.to(DEV) if torch.is_tensor(v) else v) for n, v in kw.items()}
    out = ops.resblock_pair(xa.to(DEV), *packs[DEV], k, dil, lens=lens.to(DEV), **kw_g)
    _close(out, ref, 1.2e-2 if dt == torch.bfloat16 else 2e-3, f"pair {shape} {variant}")
    assert torch.count_nonzero(out.cpu()[0, lens[0]:]) == 0 and torch.count_nonzero(out.cpu()[-1, lens[-1]:]) == 0


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 2500, 256, 512, 3, 1), (3, 1700, 512, 256, 5, 2), (1, 9000, 128, 1024, 1, 1),
                                   (5, 1153, 320, 768, 3, 1)])
def test_conv_igemm_2cta_pairs(shape, dt):
    """Shapes large enough for 256-wide tiles: the cta_group::2 kernel (CTA pairs sharing the weight tile),
    including odd numbers of M-tiles (padding tile), ragged lengths, residuals and both output kinds."""
    B, T, Cin, Cout, k, dil = shape
    torch.manual_seed(5)
    x = (torch.randn(B, T, Cin) * 0.5).to(dt)
    w = torch.randn(k, Cout, Cin) / (Cin * k) ** 0.5
    bias = torch.randn(Cout)
    r1 = torch.randn(B, T, Cout)
    lens = _lens(B, T)
    pw_c = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, "cpu")
    pw_g = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, DEV)
    ref_raw, ref_act = sim.conv(x, pw_c, res1=r1, scale=0.7, raw=torch.float32, act_out=dt, act=ops.ACT_LRELU, slope=0.2, lens=lens)
    raw, act = ops.conv(x.to(DEV), pw_g, res1=r1.to(DEV), scale=0.7, raw=torch.float32, act_out=dt, act=ops.ACT_LRELU,
                        slope=0.2, lens=lens.to(DEV))
    _close(raw, ref_raw, 2e-4, "raw")
    _close(act, ref_act, 1e-2 if dt == torch.bfloat16 else 2e-3, "act")
    # all-16-bit outputs take the TMA epilogue
    ref16, _ = sim.conv(x, pw_c, res1=r1.to(dt), raw=dt, lens=lens)
    out16, _ = ops.conv(x.to(DEV), pw_g, res1=r1.to(dt).to(DEV), raw=dt, lens=lens.to(DEV))
    _close(out16, ref16, 1e-2 if dt == torch.bfloat16 else 2e-3, "raw16")


@pytest.mark.parametrize("shape", [(2, 300, 256, 520, 3, 1), (3, 95, 512, 648, 1, 1), (16, 150, 512, 512, 5, 1),
                                   (1, 700, 1216, 1024, 3, 1), (4, 333, 192, 392, 3, 2)])
@pytest.mark.parametrize("kinds", ["res+raw+act", "raw", "res+act", "res+raw"])
def test_conv_fp32_tma_epilogue(shape, kinds):
    """fp32 residual in / fp32 raw out (+ 16-bit activated out) on wide layers: the TMA epilogue with 32-channel fp32
    tiles (epi_tma = 2).  Output widths that are not multiples of 16 / 32 / the tile, small problems (128-wide tiles),
    ragged lengths, a scale, and strided (channel-slice) outputs."""
    B, T, Cin, Cout, k, dil = shape
    torch.manual_seed(17)
    dt = torch.float16
    x = (torch.randn(B, T, Cin) * 0.5).to(dt)
    w = torch.randn(k, Cout, Cin) / (Cin * k) ** 0.5
    bias = torch.randn(Cout)
    r1 = torch.randn(B, T, Cout) if "res" in kinds else None
    lens = _lens(B, T)
    pw_c = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, "cpu")
    pw_g = ops.pack_conv(w, bias, ops.taps_1d(k, dil), dt, DEV)
    want_raw = torch.float32 if "raw" in kinds else None
    want_act = dt if "act" in kinds else None
    ref_raw, ref_act = sim.conv(x, pw_c, res1=r1, scale=0.7, raw=want_raw, act_out=want_act, act=ops.ACT_LRELU, slope=0.2, lens=lens)
    # outputs as channel slices of wider buffers (row stride > Cout), as the decoder's concatenated inputs are written
    wide_raw = torch.full((B, T, Cout + 24), 7.0, device=DEV) if want_raw is not None else None
    wide_act = torch.full((B, T, Cout + 40), 7.0, device=DEV, dtype=dt) if want_act is not None else None
    raw, act = ops.conv(x.to(DEV), pw_g, res1=None if r1 is None else r1.to(DEV), scale=0.7,
                        raw=None if wide_raw is None else wide_raw[..., 8:8 + Cout],
                        act_out=None if wide_act is None else wide_act[..., 16:16 + Cout],
                        act=ops.ACT_LRELU, slope=0.2, lens=lens.to(DEV))
    if want_raw is not None:
        _close(raw, ref_raw, 2e-4, "raw")
        assert float(wide_raw[..., :8].min()) == 7.0 and float(wide_raw[..., 8 + Cout:].min()) == 7.0   # neighbours untouched
        for b in range(B):
            assert torch.count_nonzero(raw[b, int(lens[b]):]) == 0
    if want_act is not None:
        _close(act, ref_act, 2e-3, "act")
        assert float(wide_act[..., :16].float().min()) == 7.0 and float(wide_act[..., 16 + Cout:].float().min()) == 7.0


def test_conv_igemm_2cta_full_ring_two_streams():
    """The CTA-pair implicit GEMM with its ring running full: an epilogue-bound launch (fp32 residual in, fp32 raw + f16
    activated out: acoustic AdainResBlk convs) repeated back to back on one stream while a second stream keeps the GPU
    busy.  With a 6-stage ring this raised a sticky 'illegal memory access' within ~500 launches;
    the ring is capped at 4 stages.  6000 launches, then the result against the fp32 model."""
    torch.manual_seed(21)
    B, T, C = 16, 800, 512
    x = torch.randn(B, T, C).to(torch.float16)
    w = torch.randn(3, C, C) / (3 * C) ** 0.5
    bias = torch.randn(C) * 0.1
    res = torch.randn(B, T, C)
    lens = torch.full((B,), T, dtype=torch.int32)
    packs = {d: ops.pack_conv(w, bias, ops.taps_1d(3, 1), torch.float16, d) for d in ("cpu", DEV)}
    ref_raw, ref_act = sim.conv(x, packs["cpu"], res1=res, raw=torch.float32, act_out=torch.float16, act=ops.ACT_LRELU,
                                slope=0.2, lens=lens)
    xg, rg, lg = x.to(DEV), res.to(DEV), lens.to(DEV)
    raw = torch.empty(B, T, C, device=DEV)
    act = torch.empty(B, T, C, device=DEV, dtype=torch.float16)
    big = torch.randn(32 << 20, device=DEV)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for i in range(1500):
        with torch.cuda.stream(sa):
            big.mul_(1.0)
        with torch.cuda.stream(sb):
            for _ in range(4):
                ops.conv(xg, packs[DEV], res1=rg, raw=raw, act_out=act, act=ops.ACT_LRELU, slope=0.2, lens=lg)
        if i % 100 == 99:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    _close(raw, ref_raw, 2e-3)
    _close(act, ref_act, 4e-3)


@pytest.mark.parametrize("shape", [(2, 70, 80, 64, 64, 3, 3), (1, 33, 40, 128, 128, 3, 3), (2, 50, 10, 64, 128, 3, 3),
                                   (1, 15, 5, 512, 512, 5, 5)])
def test_conv_igemm_2d(shape):
    B, T, F, Cin, Cout, kt, kf = shape
    dt = torch.float16
    torch.manual_seed(2)
    x = (torch.randn(B, T, F, Cin) * 0.5).to(dt)
    w = torch.randn(kt * kf, Cout, Cin) / (Cin * kt * kf) ** 0.5
    if kt == 5:   # 'valid' 5x5 (models.py:391)
        taps, oshape = ops.taps_2d(5, 5, 0, 0), (T - 4, F - 4)
    else:
        taps, oshape = ops.taps_2d(3, 3, 1, 1), None
    ref, _ = sim.conv(x, ops.pack_conv(w, None, taps, dt, "cpu"), out_shape=oshape, raw=torch.float32)
    out, _ = ops.conv(x.to(DEV), ops.pack_conv(w, None, taps, dt, DEV), out_shape=oshape, raw=torch.float32)
    _close(out, ref, 2e-4)


def test_conv_channel_slice_views():
    """Concat fusion: read a channel slice of a wide buffer, write into a slice of another."""
    dt = torch.float16
    torch.manual_seed(3)
    big = (torch.randn(2, 90, 1216) * 0.5).to(dt)
    w = torch.randn(3, 64, 512) / 40
    pw_c = ops.pack_conv(w, None, ops.taps_1d(3), dt, "cpu")
    pw_g = ops.pack_conv(w, None, ops.taps_1d(3), dt, DEV)
    ref, _ = sim.conv(big[..., 512:1024], pw_c, raw=torch.float32)
    dst = torch.zeros(2, 90, 256, dtype=dt, device=DEV)
    ops.conv(big.to(DEV)[..., 512:1024], pw_g, raw=dst[..., 128:192])
    _close(dst[..., 128:192], ref, 2e-3)
    assert dst[..., :128].abs().max().item() == 0 and dst[..., 192:].abs().max().item() == 0


def test_embed_layernorm():
    torch.manual_seed(4)
    tok = torch.randint(0, 178, (3, 57))
    table = torch.randn(178, 512)
    lens = _lens(3, 57)
    r32, r16 = sim.embed(tok, table, 22.6, lens, out32=True, out16=torch.float16)
    o32, o16 = ops.embed(tok.to(DEV), table.to(DEV), 22.6, lens.to(DEV), out32=True, out16=torch.float16)
    _close(o32, r32, 1e-6); _close(o16, r16, 1e-3)
    for C in (512, 256, 1024):
        x = torch.randn(3, 57, C) * 3 + 1
        g, b = torch.randn(C), torch.randn(C)
        ra, rb = sim.layernorm(x, g, b, 1e-4, act=ops.ACT_RELU, lens=lens, out_a=torch.float32, out_b=torch.float16)
        oa, ob = ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), 1e-4, act=ops.ACT_RELU, lens=lens.to(DEV),
                               out_a=torch.float32, out_b=torch.float16)
        _close(oa, ra, 1e-5); _close(ob, rb, 1e-3)
    # channel counts / dtypes outside the 16-byte fast path take the scalar kernel
    for C, dt in ((192, torch.float32), (512, torch.float16)):
        x = (torch.randn(3, 57, C) * 3 + 1).to(dt)
        g, b = torch.randn(C), torch.randn(C)
        ra, _ = sim.layernorm(x, g, b, 1e-5, lens=lens, out_a=torch.float32)
        oa, _ = ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), 1e-5, lens=lens.to(DEV), out_a=torch.float32)
        _close(oa, ra, 1e-4)


@pytest.mark.parametrize("out_dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("T", [37, 64, 150, 333, 928])
def test_relpos_attention(T, out_dt):
    """Tensor-core rel-pos attention (mma.sync, fp16 operands, fp32 accumulate / softmax) against the fp32 model of
    RelTransformerEnc.py:138-169.  Tolerance = fp16 operand rounding (2^-11 relative on q, k, v and p): the end-to-end
    mel tolerance of north_star is checked on top of it by tests/test_e2e_gpu.py."""
    torch.manual_seed(5)
    B, H, D = 2, 4, 128
    qkv = torch.randn(B, T, 3 * H * D)
    ek, ev = torch.randn(1, 9, D) * D ** -0.5, torch.randn(1, 9, D) * D ** -0.5
    lens = _lens(B, T)
    ref = sim.relpos_attention(qkv, ek, ev, 4, H, lens, torch.float32)
    out = ops.relpos_attention(qkv.to(DEV), ek[0].contiguous().to(DEV), ev[0].contiguous().to(DEV), 4, H, lens.to(DEV), out_dt)
    assert out.dtype == out_dt
    _close(out, ref, 3e-3 if out_dt == torch.float32 else 4e-3)
    for b in range(B):                         # rows beyond an utterance's length are exact zeros
        assert float(out[b, int(lens[b]):].abs().max()) == 0.0 if int(lens[b]) < T else True
    # a strided qkv view (row stride > 3*H*D) and peaked scores (large |q.k|: the online softmax must not overflow)
    wide = torch.randn(B, T, 3 * H * D + 64)
    wide[..., :H * D] *= 6.0
    ref2 = sim.relpos_attention(wide[..., :3 * H * D].contiguous(), ek, ev, 4, H, lens, torch.float32)
    out2 = ops.relpos_attention(wide.to(DEV)[..., :3 * H * D], ek[0].contiguous().to(DEV), ev[0].contiguous().to(DEV), 4, H,
                                lens.to(DEV), out_dt)
    assert bool(torch.isfinite(out2).all())
    _close(out2, ref2, 2e-2)


@pytest.mark.parametrize("T", [50, 240, 431, 600])
def test_conformer_attention(T):
    """T <= 448: tensor-core kernel (mma.sync, fp16 operands, fp32 accumulate / softmax: tolerance = fp16 operand
    rounding); longer sequences keep the fp32 CUDA-core kernel (position scores no longer fit in shared memory)."""
    torch.manual_seed(6)
    B, H, D = 2, 4, 64
    qkv = torch.randn(B, T, 3 * H * D)
    pos = torch.randn(T, H * D)
    u, v = torch.randn(H, D) * 0.1, torch.randn(H, D) * 0.1
    lens = _lens(B, T)
    q, k, vv = qkv[..., :256], qkv[..., 256:512], qkv[..., 512:]
    ref = sim.conformer_attention(q, k, vv, pos, u, v, H, lens, torch.float32)
    g = qkv.to(DEV)
    out = ops.conformer_attention(g[..., :256], g[..., 256:512], g[..., 512:], pos.to(DEV), u.to(DEV), v.to(DEV),
                                  H, lens.to(DEV), torch.float32)
    _close(out, ref, 3e-3 if T <= 448 else 2e-5)
    for b in range(B):
        assert float(out[b, int(lens[b]):].abs().max()) == 0.0 if int(lens[b]) < T else True
    out16 = ops.conformer_attention(g[..., :256], g[..., 256:512], g[..., 512:], pos.to(DEV), u.to(DEV), v.to(DEV),
                                    H, lens.to(DEV), torch.float16)
    _close(out16, ref, 4e-3)


@pytest.mark.parametrize("up", [False, True])
def test_instnorm_adain(up):
    torch.manual_seed(7)
    B, T, C = 3, 91, 640
    x = (torch.randn(B, T, C) * 2 + 0.5).half()
    lens = _lens(B, T)
    gb = torch.randn(B, 3000)[:, 100:100 + 2 * C]
    st_ref = sim.instnorm_stats(x, lens)
    st = ops.instnorm_stats(x.to(DEV), lens.to(DEV))
    _close(st, st_ref, 1e-4, "stats")
    up_w = torch.randn(C, 1, 3) if up else None
    up_b = torch.randn(C) if up else None
    ref = sim.adain_apply(x, st_ref, gb, 0.2, lens, torch.float16, up_w, up_b)
    gbg = torch.randn(B, 3000).to(DEV)
    gbg[:, 100:100 + 2 * C] = gb.to(DEV)
    out = ops.adain_apply(x.to(DEV), st, gbg[:, 100:100 + 2 * C], 0.2, lens.to(DEV), torch.float16,
                          None if not up else up_w.to(DEV).contiguous(), None if not up else up_b.to(DEV))
    _close(out, ref, 2e-3, "adain")


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("shape", [(3, 100, 96), (2, 801, 512), (1, 1600, 64), (24, 200, 1024), (10, 800, 1216), (2, 1700, 256), (16, 800, 1024),
                                   (4, 300, 384)])
def test_adain_norm_fused(shape, up):
    """as_adain_norm_apply (single pass, slab in shared memory; fp32 inputs of up to ~860 frames take the
    persistent cp.async ring kernel with 2, 3 or 4 stages -- the large shapes make it wrap the ring -- longer ones the
    single-slab TMA kernel, then the cluster / two-pass paths)."""
    B, T, C = shape
    torch.manual_seed(3)
    x = torch.randn(B, T, C) * 0.7 + torch.randn(C) * 30.0       # bias-dominated channels (|mean| >> std)
    gb = torch.randn(B, 2 * C) * 0.3
    lens = _lens(B, T)
    up_w = torch.randn(C, 3) if up else None
    up_b = torch.randn(C) if up else None
    ref = sim.adain_norm(x, gb, 0.2, lens, torch.float16, up_w, up_b)
    out = ops.adain_norm(x.to(DEV), gb.to(DEV), 0.2, lens.to(DEV), torch.float16,
                         None if up_w is None else up_w.to(DEV).contiguous(), None if up_b is None else up_b.to(DEV))
    _close(out, ref, 3e-3, f"adain_norm {shape} up={up}")


def test_repeat_and_length_regulate():
    torch.manual_seed(8)
    B, Tt, C = 3, 41, 512
    x = torch.randn(B, Tt, C)
    lens_t = torch.tensor([41, 17, 30], dtype=torch.int32)
    dur = torch.randint(1, 6, (B, Tt), dtype=torch.int32)
    To = int(2 * dur.sum(1).max())
    ref, rl = sim.length_regulate(x, dur, lens_t, 2, To, out_dtype=torch.float16)
    out, ol = ops.length_regulate(x.to(DEV), dur.to(DEV), lens_t.to(DEV), 2, To, out_dtype=torch.float16)
    _close(out, ref, 1e-3); assert torch.equal(ol.cpu(), rl)
    ref = sim.repeat_rows(x, 2, lens_t, torch.float16)
    out = ops.repeat_rows(x.to(DEV), 2, lens_t.to(DEV), torch.float16)
    _close(out, ref, 1e-3)


def test_conv_small_dwconv_pools():
    torch.manual_seed(9)
    x = torch.randn(2, 45, 80, 1)
    w = torch.randn(9, 64, 1)
    sc_c = ops.pack_small_conv(w, torch.randn(64), ops.taps_2d(3, 3, 1, 1), "cpu")
    sc_g = ops.pack_small_conv(sc_c.w, sc_c.bias, sc_c.taps, DEV)
    rr, ra = sim.conv_small(x, sc_c, raw=torch.float16, act_out=torch.float16, act=ops.ACT_LRELU, slope=0.2)
    orr, oa = ops.conv_small(x.to(DEV), sc_g, raw=torch.float16, act_out=torch.float16, act=ops.ACT_LRELU, slope=0.2)
    _close(orr, rr, 2e-3); _close(oa, ra, 2e-3)
    # 1-D, Cin = 10 (EMA_conv, models.py:482)
    x1 = torch.randn(2, 77, 10)
    sc_c = ops.pack_small_conv(torch.randn(1, 64, 10), torch.randn(64), [(0, 0)], "cpu")
    sc_g = ops.pack_small_conv(sc_c.w, sc_c.bias, sc_c.taps, DEV)
    rr, _ = sim.conv_small(x1, sc_c, raw=torch.float32)
    orr, _ = ops.conv_small(x1.to(DEV), sc_g, raw=torch.float32)
    _close(orr, rr, 1e-5)
    # depthwise 3x3 stride 2 (LearnedDownSample 'half'), (1,3) stride (1,2), 1-D k3 s2, conformer k31 + GLU
    y = torch.randn(2, 45, 80, 64).half()
    for k, s, p in (((3, 3), (2, 2), (1, 1)), ((3, 1), (2, 1), (1, 0))):
        wd, bd = torch.randn(k[0] * k[1], 64), torch.randn(64)
        ref = sim.dwconv(y, wd, bd, k, s, p, act=ops.ACT_LRELU, slope=0.2, out_dtype=torch.float16)
        out = ops.dwconv(y.to(DEV), wd.to(DEV), bd.to(DEV), k, s, p, act=ops.ACT_LRELU, slope=0.2, out_dtype=torch.float16)
        _close(out, ref, 2e-3, f"dwconv {k}")
    z = torch.randn(2, 100, 512)
    wd, bd = torch.randn(31, 256) * 0.2, torch.randn(256)
    ref = sim.dwconv(z, wd, bd, (31, 1), (1, 1), (15, 0), glu=True, act=ops.ACT_SWISH, out_dtype=torch.float16)
    out = ops.dwconv(z.to(DEV), wd.to(DEV), bd.to(DEV), (31, 1), (1, 1), (15, 0), glu=True, act=ops.ACT_SWISH,
                     out_dtype=torch.float16)
    _close(out, ref, 2e-3, "glu dwconv")
    for pt, pf in ((2, 2), (2, 1)):
        ref = sim.avgpool(y, pt, pf, torch.float16)
        out = ops.avgpool(y.to(DEV), pt, pf, torch.float16)
        _close(out, ref, 2e-3, "avgpool")
    scale, shift = torch.rand(64) + 0.5, torch.randn(64)
    ref = sim.affine_act_maxpool(y, scale, shift, 0.01, 2, torch.float16)
    out = ops.affine_act_maxpool(y.to(DEV), scale.to(DEV), shift.to(DEV), 0.01, 2, torch.float16)
    _close(out, ref, 2e-3, "maxpool")
    # fp32 tensors / odd channel counts take the scalar kernels (the fp16 cases above take the 8-channel ones)
    yf = torch.randn(2, 21, 10, 12)
    wd, bd = torch.randn(9, 12), torch.randn(12)
    _close(ops.dwconv(yf.to(DEV), wd.to(DEV), bd.to(DEV), (3, 3), (2, 2), (1, 1), out_dtype=torch.float32),
           sim.dwconv(yf, wd, bd, (3, 3), (2, 2), (1, 1), out_dtype=torch.float32), 1e-5, "dwconv fp32")
    _close(ops.avgpool(yf.to(DEV), 2, 2, torch.float32), sim.avgpool(yf, 2, 2, torch.float32), 1e-5, "avgpool fp32")
    sc12, sh12 = torch.rand(12) + 0.5, torch.randn(12)
    _close(ops.affine_act_maxpool(yf.to(DEV), sc12.to(DEV), sh12.to(DEV), 0.01, 2, torch.float32),
           sim.affine_act_maxpool(yf, sc12, sh12, 0.01, 2, torch.float32), 1e-5, "maxpool fp32")
    for ts in (1, 2):
        ref = sim.global_avgpool(y, 0.2, torch.float32, ts)
        out = ops.global_avgpool(y.to(DEV), 0.2, torch.float32, ts)
        _close(out, ref, 1e-4, "gap")


@pytest.mark.parametrize("B,T", [(5, 60), (16, 240), (19, 33)])
@pytest.mark.parametrize("H", [128, 256])
def test_bilstm(H, B, T):
    torch.manual_seed(10)
    xp = torch.randn(B, T, 8 * H)
    whh_t = torch.randn(2, H, 4 * H) / H ** 0.5
    lens = _lens(B, T)
    ref = sim.bilstm(xp, whh_t, H, lens, torch.float32)
    out = ops.bilstm(xp.to(DEV), whh_t.to(DEV), H, lens.to(DEV), torch.float32)
    _close(out, ref, 1e-3)   # fp16 tensor-core recurrence (rounding points differ from the model's)
    ref = sim.lstm_onestep(xp, H, torch.float32)
    out = ops.lstm_onestep(xp.to(DEV), H, torch.float32)
    _close(out, ref, 1e-5)


def test_layout_and_energy():
    torch.manual_seed(11)
    mel = torch.randn(3, 80, 123)
    lens = _lens(3, 123)
    _close(ops.log_norm(mel.to(DEV)), sim.log_norm(mel), 1e-5)
    ref = sim.to_channels_last(mel, torch.float16, lens)
    out = ops.to_channels_last(mel.to(DEV), torch.float16, lens.to(DEV))
    _close(out, ref, 1e-3)
    sub, mul = torch.randn(80), torch.rand(80) + 0.5
    cl = torch.randn(3, 123, 80)
    ref = sim.to_channels_first(cl, torch.float32, lens, sub, mul)
    out = ops.to_channels_first(cl.to(DEV), torch.float32, lens.to(DEV), sub.to(DEV), mul.to(DEV))
    _close(out, ref, 1e-5)
