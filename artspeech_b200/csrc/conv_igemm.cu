// Implicit-GEMM convolution for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> fused
// epilogue.  One kernel family serves every dense contraction on the synthesis path
// (SURVEY.md §10): Conv1d / polyphase ConvTranspose1d / Conv2d / Linear.
//
// GEMM view (per output tile): D[128 rows, BN] += sum over (tap j, channel chunk c) of
//     A_j,c[128, BK] * W_j,c[BN, BK]^T
// rows   = a tT x tF rectangle of output positions of one batch item (tT*tF = 128),
// A_j,c  = the same rectangle of the channels-last input, shifted by the tap offset (dt, df):
//          ONE 4-D TMA box per (tap, chunk); out-of-range coordinates are zero-filled by the TMA
//          unit, which is exactly the convolution's zero padding,
// W_j,c  = a [BN, BK] K-major box of the packed weight matrix [ntaps*CoutP, CinP].
// Both operands use the canonical K-major 128B (BK=64) / 64B (BK=32) swizzled layout that TMA
// writes and the UMMA shared-memory descriptor reads.  fp32 accumulators live in TMEM.
//
// Persistent CTAs (grid = #SMs x CTAs/SM) walk a static tile schedule.  Warp roles (320 threads):
// warp 0 = TMA producer (one lane, runs ahead across tiles through the smem ring), warp 1 = TMEM
// allocator + MMA issuer (one lane, alternates between two TMEM accumulator stages), warps 2..9 =
// two epilogue warpgroups that drain the accumulator stages alternately (TMEM lane quadrant =
// warp_idx % 4), so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv_igemm_impl.cuh"

namespace asb {

// conv_cout1.cu: single-output-channel 1-D convolutions are an HBM stream, not a GEMM
bool conv_cout1_eligible(const as_conv_params* p);
int conv_cout1_launch(const as_conv_params* p, cudaStream_t st);

// instantiated in conv_inst_*.cu
extern template int launch_conv_2cta<64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, cudaStream_t);
extern template int launch_conv_2cta<64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, cudaStream_t);
extern template int launch_conv<16, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<16, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<32, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<32, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<64, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<64, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<128, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<128, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<256, 32, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<256, 32, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<16, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<16, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<32, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<32, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<64, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<64, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<128, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<128, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<256, 64, true>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_conv<256, 64, false>(const CUtensorMap&, const CUtensorMap&, const EpiMaps&, ConvArgs&, int, cudaStream_t);
extern template int launch_halo_sw<16, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<16, 32, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<16, 64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<16, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<32, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<32, 32, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<32, 64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<32, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<64, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<64, 32, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<64, 64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<64, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<128, 32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<128, 32, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<128, 64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo_sw<128, 64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<16, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<16, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<32, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<32, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<64, true>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);
extern template int launch_halo<64, false>(const as_conv_params*, ConvArgs&, const EpiMaps&, EncodeTiledFn, cudaStream_t);

static bool halo_sw_eligible(const as_conv_params* p, int bn, size_t extra_smem) {
  if (p->F != 1 || p->Fo != 1 || p->To != p->T) return false;
  if (!(p->Cin == 32 || p->Cin == 64 || p->Cin == 128) || p->CinP != p->Cin || bn > 128 || p->CoutP != bn) return false;
  int lo = 0, hi = 0;
  for (int j = 0; j < p->ntaps; ++j) {
    if (p->tap_df[j] != 0) return false;
    lo = p->tap_dt[j] < lo ? p->tap_dt[j] : lo; hi = p->tap_dt[j] > hi ? p->tap_dt[j] : hi;
  }
  if (128 + hi - lo > 248) return false;
  const size_t HRP = (128 + hi - lo + 7) / 8 * 8;
  const size_t a_bytes = (size_t)HRP * p->Cin * 2, w_bytes = (size_t)p->ntaps * p->Cin * bn * 2;
  return 2 * a_bytes + w_bytes + extra_smem + 16 * 1024 <= 216 * 1024;
}

static bool halo_eligible(const as_conv_params* p, int bn) {
  if (p->F != 1 || p->Fo != 1 || p->To != p->T) return false;
  // measured on B200 (tools/prof_conv.py): a win for 32 channels (-15 %), a loss for 64 (the wider
  // kernel keeps two CTAs per SM there, which hides the epilogue's residual-load latency better)
  if (p->Cin % 16 != 0 || p->Cin > 32 || p->CinP != p->Cin || bn > 64 || p->CoutP != bn) return false;
  int lo = 0, hi = 0;
  for (int j = 0; j < p->ntaps; ++j) {
    if (p->tap_df[j] != 0) return false;
    lo = p->tap_dt[j] < lo ? p->tap_dt[j] : lo; hi = p->tap_dt[j] > hi ? p->tap_dt[j] : hi;
  }
  if (128 + hi - lo > 248) return false;
  const size_t w_bytes = (size_t)p->ntaps * (p->Cin / 8) * bn * 16;
  return w_bytes <= 100 * 1024;
}

static int esize(int dtype) { return dtype == AS_F32 ? 4 : 2; }
static int vec_ok(const void* p, long long ld, int dtype) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * esize(dtype)) & 15) == 0) ? 1 : 0;
}

}  // namespace asb

extern "C" int32_t as_conv_tile_n(int32_t Cout) { return asb::pick_tile_n(Cout); }

extern "C" int as_conv_igemm(const as_conv_params* p, void* stream) {
  using namespace asb;
  ASB_REQUIRE(p != nullptr, AS_ERR_SHAPE, "as_conv_igemm: null params");
  ASB_REQUIRE(p->x_dtype == AS_F16 || p->x_dtype == AS_BF16, AS_ERR_DTYPE,
              "as_conv_igemm: x must be f16 or bf16");
  ASB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->Cin > 0 && p->Cout > 0 && p->To > 0 && p->Fo > 0,
              AS_ERR_SHAPE, "as_conv_igemm: non-positive dimension");
  ASB_REQUIRE(p->ntaps >= 1 && p->ntaps <= CV_MAX_TAPS, AS_ERR_SHAPE,
              "as_conv_igemm: ntaps=%d out of range [1,%d]", p->ntaps, CV_MAX_TAPS);
  ASB_REQUIRE(p->x && p->w && p->tap_dt && p->tap_df, AS_ERR_SHAPE, "as_conv_igemm: null pointer");
  const int bk = p->Cin >= 64 ? 64 : 32;
  int bn = pick_tile_n(p->Cout);
  {
    // small problems: with 128 x 256 tiles a GEMM of a few thousand rows fills a fraction of the SMs and
    // every CTA streams the whole weight matrix; 128 x 128 tiles double the CTAs and halve the bytes each
    // one pulls from L2 (two CTAs per SM).  Only when the packed width allows it (CoutP % 128 == 0).
    static const int small_tiles = getenv("ASB_BN128_MAX_TILES") ? atoi(getenv("ASB_BN128_MAX_TILES")) : 74;
    if (bn == 256 && p->CoutP % 128 == 0) {
      int tF0 = 1;
      while (tF0 < 128 && (p->Fo % (tF0 * 2)) == 0) tF0 <<= 1;
      const long long m_tiles0 = (long long)p->B * ((p->To + 128 / tF0 - 1) / (128 / tF0)) * ((p->Fo + tF0 - 1) / tF0);
      if (m_tiles0 * (p->CoutP / 256) <= small_tiles) bn = 128;
    }
  }
  ASB_REQUIRE(p->CinP % bk == 0 && p->CinP >= p->Cin, AS_ERR_SHAPE,
              "as_conv_igemm: CinP=%d must be a multiple of %d and >= Cin=%d", p->CinP, bk, p->Cin);
  ASB_REQUIRE(p->CoutP % bn == 0 && p->CoutP >= p->Cout, AS_ERR_SHAPE,
              "as_conv_igemm: CoutP=%d must be a multiple of %d and >= Cout=%d", p->CoutP, bn,
              p->Cout);
  ASB_REQUIRE((p->x_ld % 8) == 0 && p->x_ld >= p->Cin, AS_ERR_ALIGN,
              "as_conv_igemm: x_ld=%lld must be a multiple of 8 and >= Cin", (long long)p->x_ld);
  ASB_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              AS_ERR_ALIGN, "as_conv_igemm: x and w must be 16-byte aligned");
  ASB_REQUIRE(p->y_raw || p->y_act, AS_ERR_SHAPE, "as_conv_igemm: no output requested");
  ASB_REQUIRE(p->stats == nullptr, AS_ERR_SHAPE, "as_conv_igemm: fused stats not available yet");
  int rc = check_arch();
  if (rc != AS_OK) return rc;
  static const bool no_cout1 = getenv("ASB_NO_COUT1") != nullptr;
  if (!no_cout1 && conv_cout1_eligible(p)) return conv_cout1_launch(p, reinterpret_cast<cudaStream_t>(stream));
  ASB_REQUIRE(!(p->y_act && p->y_act_dtype == AS_PCM16) && !(p->y_raw && p->y_raw_dtype == AS_PCM16), AS_ERR_DTYPE,
              "as_conv_igemm: AS_PCM16 output is only available for single-output-channel 1-D convolutions");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return AS_ERR_CUDA;

  // rows of a tile: tF = largest power of two dividing Fo (capped at 128), tT = 128 / tF, so the
  // 128-row rectangle never straddles the F edge (F = 80/40/20/10/5 on this path, 1 for 1-D)
  int tF = 1;
  while (tF < 128 && (p->Fo % (tF * 2)) == 0) tF <<= 1;
  const int tT = 128 / tF;

  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p->B; a.To = p->To; a.Fo = p->Fo; a.Cout = p->Cout; a.CoutP = p->CoutP;
  a.tT = tT; a.tF = tF;
  a.n_ttiles = (p->To + tT - 1) / tT;
  a.n_ftiles = (p->Fo + tF - 1) / tF;
  a.ntaps = p->ntaps; a.kchunks = p->CinP / bk;
  for (int j = 0; j < p->ntaps; ++j) { a.tap_dt[j] = p->tap_dt[j]; a.tap_df[j] = p->tap_df[j]; }
  const uint32_t fmt = (p->x_dtype == AS_BF16) ? 1u : 0u;
  a.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  a.bias = p->bias;
  a.res1 = p->res1; a.res1_dtype = p->res1_dtype; a.res1_ld = p->res1_ld;
  a.res2 = p->res2; a.res2_dtype = p->res2_dtype; a.res2_ld = p->res2_ld;
  a.out_scale = p->out_scale;
  a.y_raw = p->y_raw; a.y_raw_dtype = p->y_raw_dtype; a.y_raw_ld = p->y_raw_ld;
  a.y_act = p->y_act; a.y_act_dtype = p->y_act_dtype; a.y_act_ld = p->y_act_ld;
  a.y_raw_vec = (p->y_raw ? vec_ok(p->y_raw, p->y_raw_ld, p->y_raw_dtype) : 0) |
                ((p->res1 ? vec_ok(p->res1, p->res1_ld, p->res1_dtype) : 0) << 8) |
                ((p->res2 ? vec_ok(p->res2, p->res2_ld, p->res2_dtype) : 0) << 9);
  a.y_act_vec = p->y_act ? vec_ok(p->y_act, p->y_act_ld, p->y_act_dtype) : 0;
  a.act = p->act; a.slope = p->slope;
  a.lens = p->lens;
  a.stats = p->stats;
  {
    // fast epilogue: every present tensor is fp32 or the launch's 16-bit format, 16-byte aligned rows
    auto kind = [&](const void* ptr, int dtype, long long ld) -> int {
      if (ptr == nullptr) return 0;
      if (!vec_ok(ptr, ld, dtype)) return -1;
      if (dtype == AS_F32) return 2;
      return dtype == p->x_dtype ? 1 : -1;
    };
    a.k_res1 = kind(p->res1, p->res1_dtype, p->res1_ld);
    a.k_res2 = kind(p->res2, p->res2_dtype, p->res2_ld);
    a.k_raw = kind(p->y_raw, p->y_raw_dtype, p->y_raw_ld);
    a.k_act = kind(p->y_act, p->y_act_dtype, p->y_act_ld);
    const bool bias_ok = p->bias == nullptr || (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0;
    a.fast = (a.k_res1 >= 0 && a.k_res2 >= 0 && a.k_raw >= 0 && a.k_act >= 0 && bias_ok) ? 1 : 0;
    if (!a.fast) { a.k_res1 = a.k_res2 = a.k_raw = a.k_act = 0; }
    a.act_simple = 1;
    switch (p->act) {
      case AS_ACT_NONE: a.act_slope_eff = 1.f; break;
      case AS_ACT_LRELU: a.act_slope_eff = p->slope; a.act_simple = (p->slope <= 1.f) ? 1 : 0; break;
      case AS_ACT_RELU: a.act_slope_eff = 0.f; break;
      case AS_ACT_ABS: a.act_slope_eff = -1.f; break;
      default: a.act_simple = 0; a.act_slope_eff = 1.f; break;
    }
  }

  const CUtensorMapDataType dt =
      p->x_dtype == AS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  // TMA epilogue: 1-D, every present tensor in the launch's 16-bit format, no second residual
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  static const bool no_epi_tma = getenv("ASB_NO_EPI_TMA") != nullptr;
  a.epi_tma = (!no_epi_tma && a.fast && p->F == 1 && p->Fo == 1 && a.k_res2 == 0 && a.k_res1 != 2 && a.k_raw != 2 &&
               a.k_act != 2 && (a.k_raw == 1 || a.k_act == 1)) ? 1 : 0;
  // ... or (epi_tma = 2, wide layers only: never a halo kernel) fp32 residual and / or fp32 raw output next to an optional
  // 16-bit activated output: the residual stream between AdaIN blocks.  ASB_NO_EPI_TMA32=1 keeps the direct epilogue.
  static const bool no_epi_tma32 = getenv("ASB_NO_EPI_TMA32") != nullptr;
  if (!a.epi_tma && !no_epi_tma && !no_epi_tma32 && a.fast && p->F == 1 && p->Fo == 1 && a.k_res2 == 0 && bn >= 128 &&
      (p->Cin > 128 || p->CoutP > 128) && a.k_res1 != 1 && a.k_raw != 1 && a.k_act != 2 && (a.k_res1 == 2 || a.k_raw == 2)) {
    auto mk32 = [&](CUtensorMap* m, const void* ptr, long long ld, bool f32) -> bool {
      const cuuint64_t es_b = f32 ? 4 : 2;
      cuuint64_t dims[3] = {(cuuint64_t)p->Cout, (cuuint64_t)p->To, (cuuint64_t)p->B};
      cuuint64_t strides[2] = {(cuuint64_t)ld * es_b, (cuuint64_t)ld * es_b * p->To};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t es[3] = {1, 1, 1};
      return enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt, 3, const_cast<void*>(ptr), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = true;
    if (a.k_res1 == 2) ok = ok && mk32(&em.r1, p->res1, p->res1_ld, true);
    if (a.k_raw == 2) ok = ok && mk32(&em.raw, p->y_raw, p->y_raw_ld, true);
    if (a.k_act == 1) ok = ok && mk32(&em.act, p->y_act, p->y_act_ld, false);
    if (ok) a.epi_tma = 2;
  }
  if (a.epi_tma == 1) {
    auto mk = [&](CUtensorMap* m, const void* ptr, long long ld) -> bool {
      cuuint64_t dims[3] = {(cuuint64_t)p->Cout, (cuuint64_t)p->To, (cuuint64_t)p->B};
      cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * p->To};
      cuuint32_t box[3] = {epi_cols(bn), 32, 1};
      cuuint32_t es[3] = {1, 1, 1};
      return enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 epi_cols(bn) == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = true;
    if (a.k_res1 == 1) ok = ok && mk(&em.r1, p->res1, p->res1_ld);
    if (a.k_raw == 1) ok = ok && mk(&em.raw, p->y_raw, p->y_raw_ld);
    if (a.k_act == 1) ok = ok && mk(&em.act, p->y_act, p->y_act_ld);
    if (!ok) a.epi_tma = 0;
  }
  const CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->F, (cuuint64_t)p->T, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, (cuuint64_t)p->x_ld * 2 * p->F,
                             (cuuint64_t)p->x_ld * 2 * p->F * p->T};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)tF, (cuuint32_t)tT, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, dt, 4, const_cast<void*>(p->x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p->CinP, (cuuint64_t)p->ntaps * p->CoutP};
    cuuint64_t strides[1] = {(cuuint64_t)p->CinP * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmW, dt, 2, const_cast<void*>(p->w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool no_halo = getenv("ASB_NO_HALO") != nullptr;
  static const bool no_halo_sw = getenv("ASB_NO_HALO_SW") != nullptr;
  if (!no_halo && !no_halo_sw && halo_sw_eligible(p, bn, a.epi_tma ? epi_bytes(bn) + 1024 : 0)) {
    const bool bf = p->x_dtype == AS_BF16;
#define HALOSW_CASE(BN_)                                                                          \
  if (bn == BN_) {                                                                                \
    if (p->Cin == 32) return bf ? launch_halo_sw<BN_, 32, true>(p, a, em, enc, st) : launch_halo_sw<BN_, 32, false>(p, a, em, enc, st); \
    return bf ? launch_halo_sw<BN_, 64, true>(p, a, em, enc, st) : launch_halo_sw<BN_, 64, false>(p, a, em, enc, st);                    \
  }
    HALOSW_CASE(16) HALOSW_CASE(32) HALOSW_CASE(64) HALOSW_CASE(128)
#undef HALOSW_CASE
  }
  if (!no_halo && halo_eligible(p, bn)) {
    const bool bf = p->x_dtype == AS_BF16;
#define HALO_CASE(BN_) if (bn == BN_) return bf ? launch_halo<BN_, true>(p, a, em, enc, st) : launch_halo<BN_, false>(p, a, em, enc, st);
    HALO_CASE(16) HALO_CASE(32) HALO_CASE(64)
#undef HALO_CASE
  }
  {
    // 256-wide tiles with at least one full pair of M-tiles: CTA pairs share the weight tile (cta_group::2)
    static const bool no_2cta = getenv("ASB_NO_2CTA") != nullptr;
    const int m_tiles_all = p->B * a.n_ttiles * a.n_ftiles;
    if (!no_2cta && bn == 256 && bk == 64 && m_tiles_all >= 2) {
      CUtensorMap tmWh;
      cuuint64_t dims[2] = {(cuuint64_t)p->CinP, (cuuint64_t)p->ntaps * p->CoutP};
      cuuint64_t strides[1] = {(cuuint64_t)p->CinP * 2};
      cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)(bn / 2)};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tmWh, dt, 2, const_cast<void*>(p->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(W half) failed: %d", (int)r);
      return p->x_dtype == AS_BF16 ? launch_conv_2cta<64, true>(tmA, tmWh, em, a, st) : launch_conv_2cta<64, false>(tmA, tmWh, em, a, st);
    }
  }
  const int total_tiles = p->B * a.n_ttiles * a.n_ftiles * (p->CoutP / bn);
#define CV_CASE(BN_, BK_)                                                                    \
  if (bn == BN_ && bk == BK_)                                                                \
    return p->x_dtype == AS_BF16 ? launch_conv<BN_, BK_, true>(tmA, tmW, em, a, total_tiles, st) \
                                 : launch_conv<BN_, BK_, false>(tmA, tmW, em, a, total_tiles, st);
  CV_CASE(16, 32) CV_CASE(32, 32) CV_CASE(64, 32) CV_CASE(128, 32) CV_CASE(256, 32)
  CV_CASE(16, 64) CV_CASE(32, 64) CV_CASE(64, 64) CV_CASE(128, 64) CV_CASE(256, 64)
#undef CV_CASE
  set_error("as_conv_igemm: no kernel for tile_n=%d bk=%d", bn, bk);
  return AS_ERR_SHAPE;
}
