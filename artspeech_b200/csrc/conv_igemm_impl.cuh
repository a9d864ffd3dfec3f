// Kernel and launch templates of the implicit-GEMM convolution family (see conv_igemm.cu for the
// design notes).  Included by conv_igemm.cu (dispatch) and conv_inst_*.cu (explicit instantiations,
// split over several translation units so that nvcc compiles them in parallel).
#pragma once
#include "tc_util.cuh"

namespace asb {

constexpr int CV_THREADS = 320;  // TMA warp, MMA warp, 2 epilogue warpgroups
constexpr int CV_MAX_TAPS = 32;

struct ConvArgs {
  int B, To, Fo, Cout, CoutP;
  int tT, tF, n_ttiles, n_ftiles;
  int ntaps, kchunks, stages;
  int tap_dt[CV_MAX_TAPS];
  int tap_df[CV_MAX_TAPS];
  uint32_t idesc;
  const float* bias;
  const void* res1; int res1_dtype; long long res1_ld;
  const void* res2; int res2_dtype; long long res2_ld;
  float out_scale;
  void* y_raw; int y_raw_dtype; long long y_raw_ld; int y_raw_vec;
  void* y_act; int y_act_dtype; long long y_act_ld; int y_act_vec;
  int act; float slope;
  const int* lens;
  float* stats;
  // fast-epilogue plan (host-computed): tensor kinds 0 absent / 1 launch 16-bit format / 2 fp32
  int fast, k_res1, k_res2, k_raw, k_act, act_simple;
  float act_slope_eff;
  int halo_rows, pad_lo;   // halo variants: rows of the halo tile, -min(tap_dt)
  int epi_tma;             // 1: all-16-bit 1-D epilogue through shared memory + TMA
  int halo_baseoff;        // swizzled halo: put (row & 7) into the descriptor's base-offset field
  int skip;                // 1: tiles that lie entirely beyond their item's length are left out (conv_igemm / 2cta only)
  int pad_stages;          // CTA-pair kernel experiments: unused ring-stage-sized blocks in front of the ring (ASB_2CTA_PAD)
};

struct EpiMaps { CUtensorMap r1, raw, act; };
// per epilogue warp: 2 staging tiles of [32 rows][EPC channels]; EPC = 64 (128 B rows, 128B swizzle) or,
// for tiles narrower than 64 channels, 32 (64 B rows, 64B swizzle)
__host__ __device__ constexpr uint32_t epi_cols(int bn) { return bn >= 64 ? 64u : 32u; }
__host__ __device__ constexpr uint32_t epi_warp_bytes(int bn) { return 2u * 32u * epi_cols(bn) * 2u; }
__host__ __device__ constexpr uint32_t epi_bytes(int bn) { return 8u * epi_warp_bytes(bn) + 64u; }


// ---------------------------------------------------------------------------------------------
// Tile skipping for ragged batches: an M-tile whose first time index is >= its item's length produces only zeros
// (the epilogue masks rows t >= lens[b]).  Producer, MMA issuer and epilogue evaluate the same pure function of
// (tile, lens) and leave such tiles out of the pipeline -- no ring slot, accumulator stage or barrier phase is
// consumed -- and the epilogue's first warpgroup writes the zeros.  A 2-CTA super-tile is skipped only when both
// of its M-tiles are dead.  Batches with predicted durations / mixed-length serving otherwise compute their padding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool conv_mtile_dead(const ConvArgs& a, int mt) {
  const int m = mt / a.n_ftiles;
  const int tt = m % a.n_ttiles, b = m / a.n_ttiles;
  if (b >= a.B) return true;                       // the filler tile of an odd pair
  return tt * a.tT >= __ldg(a.lens + b);
}
__device__ __forceinline__ bool conv_unit_skip(const ConvArgs& a, int unit, int cl) {
  if (!a.skip) return false;
  for (int r = 0; r < cl; ++r)
    if (!conv_mtile_dead(a, unit * cl + r)) return false;
  return true;
}
// zeros for `ncols` channels of one output row (any output dtype / alignment)
__device__ __forceinline__ void conv_zero_row(void* base, int dtype, long long off_elems, int ncols) {
  const int es = dtype == AS_F32 ? 4 : 2;
  char* p = reinterpret_cast<char*>(base) + off_elems * es;
  const int nbytes = ncols * es;
  int i = 0;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0)
    for (; i + 16 <= nbytes; i += 16) *reinterpret_cast<uint4*>(p + i) = make_uint4(0u, 0u, 0u, 0u);
  for (; i < nbytes; i += es) {
    if (es == 4) *reinterpret_cast<float*>(p + i) = 0.f;
    else *reinterpret_cast<uint16_t*>(p + i) = 0;
  }
}
// the epilogue's share of a skipped tile: this thread's row (tile row m = q * 32 + lane) of the CTA's own M-tile
__device__ __forceinline__ void conv_zero_tile_row(const ConvArgs& a, int mt, int n0, int bn, int m) {
  const int it_ = m / a.tF, if_ = m - it_ * a.tF;
  int u = mt;
  const int ft = u % a.n_ftiles; u /= a.n_ftiles;
  const int tt = u % a.n_ttiles; u /= a.n_ttiles;
  const int b = u, t = tt * a.tT + it_, f = ft * a.tF + if_;
  if (b >= a.B || t >= a.To || f >= a.Fo) return;
  const int ncols = min(bn, a.Cout - n0);
  if (ncols <= 0) return;
  const long long row = ((long long)b * a.To + t) * a.Fo + f;
  if (a.y_raw != nullptr) conv_zero_row(a.y_raw, a.y_raw_dtype, row * a.y_raw_ld + n0, ncols);
  if (a.y_act != nullptr) conv_zero_row(a.y_act, a.y_act_dtype, row * a.y_act_ld + n0, ncols);
}

// ---------------------------------------------------------------------------------------------
// epilogue store helpers: 16 consecutive channels of one row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store16(void* base, int dtype, long long off, const float (&v)[16],
                                        int nvalid, bool vec) {
  if (dtype == AS_F32) {
    float* p = reinterpret_cast<float*>(base) + off;
    if (vec && nvalid == 16) {
      float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
      for (int i = 0; i < 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < nvalid) p[i] = v[i];
    }
  } else {
    uint16_t* p = reinterpret_cast<uint16_t*>(base) + off;
    if (vec && nvalid == 16) {
      uint4* p4 = reinterpret_cast<uint4*>(p);
      p4[0] = make_uint4(pack16(v[0], v[1], dtype), pack16(v[2], v[3], dtype),
                         pack16(v[4], v[5], dtype), pack16(v[6], v[7], dtype));
      p4[1] = make_uint4(pack16(v[8], v[9], dtype), pack16(v[10], v[11], dtype),
                         pack16(v[12], v[13], dtype), pack16(v[14], v[15], dtype));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < nvalid) p[i] = to16(v[i], dtype);
    }
  }
}

__device__ __forceinline__ void add_res16(const void* base, int dtype, long long off, float (&v)[16],
                                          int nvalid, bool vec) {
  if (dtype == AS_F32) {
    const float* p = reinterpret_cast<const float*>(base) + off;
    if (vec && nvalid == 16) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = __ldg(p4 + i);
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < nvalid) v[i] += p[i];
    }
  } else {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + off;
    if (vec && nvalid == 16) {
      const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 t = __ldg(p4 + h);
        uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[8 * h + 2 * i] += from16(uint16_t(w[i] & 0xFFFF), dtype);
          v[8 * h + 2 * i + 1] += from16(uint16_t(w[i] >> 16), dtype);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < nvalid) v[i] += from16(p[i], dtype);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Fast epilogue chunk (16 channels of one row).  The generic helpers above dispatch on dtypes per
// element and cost ~500 SASS instructions per chunk (ncu: the epilogue, not the MMAs, bounded the
// kernel).  This path is specialised at compile time on the 16-bit format and hoists every
// decision out of the chunk loop: ~120 instructions per chunk.
//   kind: 0 = absent, 1 = 16-bit (the launch's operand format), 2 = fp32
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void add_res_fast(const char* p, int kind, float (&v)[16]) {
  if (kind == 1) {
    const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(p)), t1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    unpack_add<BF16>(t0.x, v[0], v[1]); unpack_add<BF16>(t0.y, v[2], v[3]);
    unpack_add<BF16>(t0.z, v[4], v[5]); unpack_add<BF16>(t0.w, v[6], v[7]);
    unpack_add<BF16>(t1.x, v[8], v[9]); unpack_add<BF16>(t1.y, v[10], v[11]);
    unpack_add<BF16>(t1.z, v[12], v[13]); unpack_add<BF16>(t1.w, v[14], v[15]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
}
template <bool BF16>
__device__ __forceinline__ void store_fast(char* p, int kind, const float (&v)[16]) {
  if (kind == 1) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7]));
    q[1] = make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]), pack2<BF16>(v[14], v[15]));
  } else {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// epilogue (shared by both kernels): two warpgroups alternate tiles; TMEM -> registers -> global
// ---------------------------------------------------------------------------------------------
template <int BN, bool BF16>
__device__ __forceinline__ void run_epilogue(const ConvArgs& a, uint32_t tmem_base, uint32_t tfull0,
                                             uint32_t tempty0, int warp, int lane, int m_tiles,
                                             int total_tiles, const float* bias_s, int tile_first = -1,
                                             int tile_step = 0, int cl = 1, int rank = 0, bool remote_tempty = false) {
  // 2-CTA kernels: `tile` counts cluster-level super-tiles (cl adjacent M-tiles x one N-tile), m_tiles is
  // the number of M units (pairs), this CTA takes M-tile unit*cl + rank; tempty0 is then the leader's
  // barrier as a shared::cluster address.
  if (tile_first < 0) { tile_first = blockIdx.x; tile_step = gridDim.x; }
  // 16-bit residual rows are fetched into registers BEFORE the accumulator wait (ncu: the epilogue
  // warps were stalled on these loads, long_scoreboard ~12 cycles/issue): up to 128 channels/row.
  // BN = 256 runs one CTA per SM (<= 204 regs/thread): 128 channels; narrower tiles run two CTAs
  // per SM (<= 102 regs/thread): 64 channels.
  constexpr int NPRE = (BN >= 256 ? 128 : (BN < 64 ? BN : 64)) / 8;   // uint4 registers of residual prefetch
  auto tfull_bar = [&](int i) { return tfull0 + 8u * i; };
  auto tempty_bar = [&](int i) { return tempty0 + 8u * i; };
    const int wg = (warp - 2) >> 2;  // 0 / 1 = accumulator stage this warpgroup drains
  const int q = warp & 3;          // TMEM lane quadrant this warp may access
  const int m = q * 32 + lane;     // tile row
  const int it_ = m / a.tF, if_ = m - it_ * a.tF;
  int lt_live = 0;
  for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
    if (conv_unit_skip(a, tile % m_tiles, cl)) {           // all padding: zeros, no accumulator stage consumed
      if (wg == 0) conv_zero_tile_row(a, (tile % m_tiles) * cl + rank, (tile / m_tiles) * BN, BN, m);
      continue;
    }
    const int lt = lt_live++;
    if ((lt & 1) != wg) continue;
    int mt = (tile % m_tiles) * cl + rank;
    const int n0 = (tile / m_tiles) * BN;
    const int ft = mt % a.n_ftiles; mt /= a.n_ftiles;
    const int tt = mt % a.n_ttiles; mt /= a.n_ttiles;
    const int b = mt;
    const int t = tt * a.tT + it_, f = ft * a.tF + if_;
    const bool row_ok = (t < a.To) && (f < a.Fo) && (b < a.B);   // b == B: the padding tile of an odd pair
    bool masked = false;
    if (a.lens != nullptr && row_ok) masked = t >= __ldg(a.lens + b);
    const long long row = ((long long)b * a.To + t) * a.Fo + f;

    // per-tile constants of the fast path (byte pointers of this thread's row, kinds, slope)
    const int ncols = min(BN, a.Cout - n0);  // valid columns in this N tile (may be <= 0)
    const bool fast = a.fast != 0;
    const int k_r1 = a.k_res1, k_r2 = a.k_res2, k_raw = a.k_raw, k_act = a.k_act;
    const char* p_r1 = k_r1 ? reinterpret_cast<const char*>(a.res1) + (row * a.res1_ld + n0) * (k_r1 == 1 ? 2 : 4) : nullptr;
    const char* p_r2 = k_r2 ? reinterpret_cast<const char*>(a.res2) + (row * a.res2_ld + n0) * (k_r2 == 1 ? 2 : 4) : nullptr;
    char* p_raw = k_raw ? reinterpret_cast<char*>(a.y_raw) + (row * a.y_raw_ld + n0) * (k_raw == 1 ? 2 : 4) : nullptr;
    char* p_act = k_act ? reinterpret_cast<char*>(a.y_act) + (row * a.y_act_ld + n0) * (k_act == 1 ? 2 : 4) : nullptr;
    const float scale = masked ? 0.f : a.out_scale;   // masked rows become exact zeros (res are finite)
    const float aslope = a.act_slope_eff;
    const float* bias_t = bias_s + n0;
    const bool pre_ok = fast && k_r1 == 1 && row_ok && !masked;
    uint4 pre[NPRE];
    if (pre_ok) {
#pragma unroll
      for (int i = 0; i < NPRE; ++i)
        if (i * 8 + 8 <= ncols) pre[i] = __ldg(reinterpret_cast<const uint4*>(p_r1) + i);
    }

    mbar_wait(tfull_bar(wg), ((uint32_t)lt >> 1) & 1u);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (uint32_t)(wg * BN) + (uint32_t(q * 32) << 16);
#pragma unroll
    for (int cc = 0; cc < BN / 16; ++cc) {
      const int c0 = cc * 16;
      if (c0 >= ncols) break;  // warp-uniform
      uint32_t r[16];
      tc_ld16(tacc + uint32_t(c0), r);
      tc_wait_ld();
      if (!row_ok) continue;
      const int co = n0 + c0;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      if (fast && c0 + 16 <= ncols) {
        {   // bias staged in shared memory once per CTA (zeros when the layer has none)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_t + c0 + 4 * i);
            v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
          }
        }
        if (pre_ok && 2 * cc + 1 < NPRE) {
          const uint4 t0 = pre[(2 * cc) < NPRE ? 2 * cc : 0], t1 = pre[(2 * cc + 1) < NPRE ? 2 * cc + 1 : 0];
          unpack_add<BF16>(t0.x, v[0], v[1]); unpack_add<BF16>(t0.y, v[2], v[3]);
          unpack_add<BF16>(t0.z, v[4], v[5]); unpack_add<BF16>(t0.w, v[6], v[7]);
          unpack_add<BF16>(t1.x, v[8], v[9]); unpack_add<BF16>(t1.y, v[10], v[11]);
          unpack_add<BF16>(t1.z, v[12], v[13]); unpack_add<BF16>(t1.w, v[14], v[15]);
        } else if (k_r1 && !masked) add_res_fast<BF16>(p_r1 + c0 * (k_r1 == 1 ? 2 : 4), k_r1, v);
        if (k_r2 && !masked) add_res_fast<BF16>(p_r2 + c0 * (k_r2 == 1 ? 2 : 4), k_r2, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= scale;
        if (k_raw) store_fast<BF16>(p_raw + c0 * (k_raw == 1 ? 2 : 4), k_raw, v);
        if (k_act) {
          if (a.act_simple) {
            // none / lrelu / relu / abs: act(v) = max(v, v * s) with s = 1 / slope / 0 / -1
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * aslope);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], a.act, a.slope);
          }
          store_fast<BF16>(p_act + c0 * (k_act == 1 ? 2 : 4), k_act, v);
        }
      } else {
        // generic path: partial chunks (Cout not a multiple of 16), unaligned or mixed-format tensors
        const int nvalid = min(16, a.Cout - co);
        if (a.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (i < nvalid) v[i] += __ldg(a.bias + co + i);
        }
        if (a.res1 != nullptr) add_res16(a.res1, a.res1_dtype, row * a.res1_ld + co, v, nvalid, false);
        if (a.res2 != nullptr) add_res16(a.res2, a.res2_dtype, row * a.res2_ld + co, v, nvalid, false);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = masked ? 0.f : v[i] * a.out_scale;
        if (a.y_raw != nullptr) store16(a.y_raw, a.y_raw_dtype, row * a.y_raw_ld + co, v, nvalid, false);
        if (a.y_act != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], a.act, a.slope);
          store16(a.y_act, a.y_act_dtype, row * a.y_act_ld + co, v, nvalid, false);
        }
      }
    }
    // all TMEM reads of this stage are complete (tcgen05.wait::ld above): hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (remote_tempty) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_bar(wg)) : "memory");
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(wg)) : "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA epilogue for the all-16-bit 1-D case (every vocoder conv).  ncu on the register->global
// epilogue above: l1tex__data_pipe_lsu_wavefronts at 68 % of peak — a thread owns a row, so every
// 16-byte access of a warp lands in a different 128-byte line (32 wavefronts per instruction).
// Here each epilogue warp stages its 32 rows x 64 channels in a 128B-swizzled shared tile and moves
// it with ONE bulk tensor copy per tensor: the residual arrives by TMA load (issued before the
// accumulator wait), raw / activated outputs leave by TMA store.  Out-of-range rows / channels are
// clipped by the TMA unit, so partial tiles need no special path.
// ---------------------------------------------------------------------------------------------
template <int BN, bool BF16, bool F32 = false>
__device__ __forceinline__ void run_epilogue_tma(const ConvArgs& a, const EpiMaps& maps, uint32_t tmem_base,
                                                 uint32_t tfull0, uint32_t tempty0, int warp, int lane, int m_tiles,
                                                 int total_tiles, const float* bias_s, uint32_t epi_base,
                                                 int tile_first = -1, int tile_step = 0, int cl = 1, int rank = 0,
                                                 bool remote_tempty = false) {
  // F32 (epi_tma == 2): the residual and / or the raw output are fp32 (the residual stream between AdaIN blocks), the
  // activated output 16-bit.  Groups of 32 channels: fp32 tiles of [32 rows][32 ch] (128 B rows, 128B swizzle), the
  // activated tile [32][32] 16-bit (64 B rows, 64B swizzle).  The direct epilogue these launches used before had every
  // lane store its own row (32 wavefronts per instruction) and ran 60-70 % slower than the MMAs.
  if (tile_first < 0) { tile_first = blockIdx.x; tile_step = gridDim.x; }
  const int ew = warp - 2;                 // 0..7
  const int wg = ew >> 2;
  const int q = warp & 3;
  constexpr uint32_t EPC = F32 ? 32u : epi_cols(BN);      // channels per staged group
  constexpr uint32_t ROWA = F32 ? 128u : EPC * 2u;        // bytes per staged row of the residual / raw tile
  constexpr uint32_t ROWB = EPC * 2u;                     // ... of the activated tile (16-bit)
  constexpr uint32_t TILEA = 32u * ROWA;
  constexpr int CPG = EPC / 16;                           // 16-channel chunks per group
  const uint32_t bufA = epi_base + (uint32_t)ew * epi_warp_bytes(BN);   // residual in / raw out (in place)
  const uint32_t bufB = bufA + epi_warp_bytes(BN) / 2u;                  // activated out
  const uint32_t rbar = epi_base + 8u * epi_warp_bytes(BN) + 8u * ew;
  const bool has_r1 = a.k_res1 != 0, has_raw = a.k_raw != 0, has_act = a.k_act != 0;
  const uint32_t rowoffA = (uint32_t)lane * ROWA, rowoffB = (uint32_t)lane * ROWB;
  // 16-byte unit swizzle of the TMA layout: 128B mode XORs with (row & 7), 64B mode with (row >> 1) & 3
  const uint32_t swA = (ROWA == 128) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
  const uint32_t swB = (ROWB == 128) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
  uint32_t rphase = 0;
  int lt_live = 0;
  for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
    if (conv_unit_skip(a, tile % m_tiles, cl)) {           // all padding: zeros (plain stores), no accumulator stage consumed
      if (wg == 0) conv_zero_tile_row(a, (tile % m_tiles) * cl + rank, (tile / m_tiles) * BN, BN, q * 32 + lane);
      continue;
    }
    const int lt = lt_live++;
    if ((lt & 1) != wg) continue;
    const int mt = (tile % m_tiles) * cl + rank;
    const int n0 = (tile / m_tiles) * BN;
    const int tt = mt % a.n_ttiles, b = mt / a.n_ttiles;      // b == B (padding tile of an odd pair): TMA clips everything
    const int t_base = tt * 128 + q * 32;
    const int t = t_base + lane;
    const bool masked = (a.lens != nullptr) && (b < a.B) && (t < a.To) && (t >= __ldg(a.lens + b));
    const int ncols = min(BN, a.Cout - n0);
    const int ngroups = (ncols + (int)EPC - 1) / (int)EPC;
    const float scale = a.out_scale, aslope = a.act_slope_eff;

    if (has_r1 && lane == 0 && ngroups > 0) {
      bulk_wait_read0();                                   // previous stores have drained this buffer
      mbar_expect_tx(rbar, TILEA);
      tma_load_3d(bufA, &maps.r1, rbar, n0, t_base, b);
    }
    mbar_wait(tfull0 + 8u * wg, ((uint32_t)lt >> 1) & 1u);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (uint32_t)(wg * BN) + (uint32_t(q * 32) << 16);
    for (int g = 0; g < ngroups; ++g) {
      if (has_r1) {
        mbar_wait(rbar, rphase);
        rphase ^= 1u;
      } else {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
#pragma unroll
      for (int cc = 0; cc < CPG; ++cc) {
        const int c0 = g * (int)EPC + cc * 16;
        if (c0 < ncols) {   // warp-uniform
          uint32_t r[16];
          tc_ld16(tacc + uint32_t(c0), r);
          tc_wait_ld();
          float v[16];
          const float* bt = bias_s + n0 + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = *reinterpret_cast<const float4*>(bt + 4 * i);
            v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
          }
          if constexpr (F32) {
            // fp32 tile: this row's 16 channels are 4 units of 16 bytes
            uint32_t oa[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) oa[j] = bufA + rowoffA + (((uint32_t)(4 * cc + j) ^ swA) << 4);
            if (has_r1) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 t4 = lds128(oa[j]);
                v[4 * j] += __uint_as_float(t4.x); v[4 * j + 1] += __uint_as_float(t4.y);
                v[4 * j + 2] += __uint_as_float(t4.z); v[4 * j + 3] += __uint_as_float(t4.w);
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = masked ? 0.f : v[i] * scale;
            if (has_raw) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(oa[j], make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                         __float_as_uint(v[4 * j + 3])));
            }
          } else {
            const uint32_t o0 = rowoffA + (((uint32_t)(2 * cc) ^ swA) << 4), o1 = rowoffA + (((uint32_t)(2 * cc + 1) ^ swA) << 4);
            if (has_r1) {
              const uint4 t0 = lds128(bufA + o0), t1 = lds128(bufA + o1);
              unpack_add<BF16>(t0.x, v[0], v[1]); unpack_add<BF16>(t0.y, v[2], v[3]);
              unpack_add<BF16>(t0.z, v[4], v[5]); unpack_add<BF16>(t0.w, v[6], v[7]);
              unpack_add<BF16>(t1.x, v[8], v[9]); unpack_add<BF16>(t1.y, v[10], v[11]);
              unpack_add<BF16>(t1.z, v[12], v[13]); unpack_add<BF16>(t1.w, v[14], v[15]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = masked ? 0.f : v[i] * scale;
            if (has_raw) {
              sts128(bufA + o0, make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7])));
              sts128(bufA + o1, make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]), pack2<BF16>(v[14], v[15])));
            }
          }
          if (has_act) {
            if (a.act_simple) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * aslope);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], a.act, a.slope);
            }
            const uint32_t b0 = bufB + rowoffB + (((uint32_t)(2 * cc) ^ swB) << 4), b1 = bufB + rowoffB + (((uint32_t)(2 * cc + 1) ^ swB) << 4);
            sts128(b0, make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7])));
            sts128(b1, make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]), pack2<BF16>(v[14], v[15])));
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (has_raw) tma_store_3d(&maps.raw, bufA, n0 + g * (int)EPC, t_base, b);
        if (has_act) tma_store_3d(&maps.act, bufB, n0 + g * (int)EPC, t_base, b);
        bulk_commit();
        if (has_r1 && g + 1 < ngroups) {
          bulk_wait_read0();
          mbar_expect_tx(rbar, TILEA);
          tma_load_3d(bufA, &maps.r1, rbar, n0 + (g + 1) * (int)EPC, t_base, b);
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (remote_tempty) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty0 + 8u * wg) : "memory");
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * wg) : "memory");
    }
  }
  if (lane == 0) bulk_wait_all0();   // all stores complete before the CTA exits
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int BK, bool BF16>
__global__ void __launch_bounds__(CV_THREADS, (BN >= 256 ? 1 : 2))
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ EpiMaps emaps,
                  const __grid_constant__ ConvArgs a) {
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr int W_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);   // two accumulator stages (power of two)

  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int S = a.stages;
  const uint32_t bar_base = smem_base + S * STAGE_BYTES;  // full[S], empty[S], tfull[2], tempty[2], slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * S + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * S + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bar_base + 8u * (2 * S + 6) - smem_u32(smem_dyn)));
  for (int i = threadIdx.x; i < a.CoutP; i += blockDim.x) bias_s[i] = (a.bias != nullptr && i < a.Cout) ? a.bias[i] : 0.f;
  const uint32_t epi_base = (bar_base + 8u * (2 * S + 6) + (uint32_t)a.CoutP * 4u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = a.B * a.n_ttiles * a.n_ftiles;
  const int total_tiles = m_tiles * (a.CoutP / BN);
  const int k_iters = a.ntaps * a.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    if (a.epi_tma) for (int i = 0; i < 8; ++i) mbar_init(epi_base + 8u * epi_warp_bytes(BN) + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // barrier init / TMEM alloc / descriptor prefetch above overlap the previous kernel's tail
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // Persistent CTA: every role walks the same static tile schedule (tile = blockIdx.x + i*gridDim.x,
  // M fastest so concurrently running CTAs share the weight tile in L2).  The smem ring and the two
  // TMEM accumulator stages decouple the roles: TMA runs ahead across tile boundaries, the MMA of
  // tile i+1 overlaps the epilogue of tile i.
  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (conv_unit_skip(a, tile % m_tiles, 1)) continue;
        int mt = tile % m_tiles;
        const int n0 = (tile / m_tiles) * BN;
        const int ft = mt % a.n_ftiles; mt /= a.n_ftiles;
        const int tt = mt % a.n_ttiles; mt /= a.n_ttiles;
        const int b = mt, t0 = tt * a.tT, f0 = ft * a.tF;
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          const int tap = kit / a.kchunks, kc = kit - tap * a.kchunks;
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          tma_load_4d(sa, &tmA, full_bar(s), kc * BK, f0 + a.tap_df[tap], t0 + a.tap_dt[tap], b);
          tma_load_2d(sa + A_BYTES, &tmW, full_bar(s), kc * BK, tap * a.CoutP + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elected lane of a converged warp: back-to-back UTCHMMA issue (no per-instruction election loop)
      // ===== MMA issuer =====
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (conv_unit_skip(a, tile % m_tiles, 1)) continue;
        const int acc = lt & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)lt >> 1) & 1u) ^ 1u);   // epilogue drained this stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint64_t da = make_smem_desc<BK>(sa);
          const uint64_t db = make_smem_desc<BK>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle row: +2 in the >>4 field
            tc_mma_f16(tacc, da + uint64_t(2 * k), db + uint64_t(2 * k), a.idesc, (kit | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
        }
        tc_commit(tfull_bar(acc));  // accumulator of this tile complete
        ++lt;
      }
    }
  } else {
    // ===== epilogue =====
    if (a.epi_tma == 2) {
      if constexpr (BN >= 128)
        run_epilogue_tma<BN, BF16, true>(a, emaps, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s, epi_base);
    } else if (a.epi_tma)
      run_epilogue_tma<BN, BF16>(a, emaps, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s, epi_base);
    else
      run_epilogue<BN, BF16>(a, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2) for 256-wide tiles.  ncu on the 1-CTA kernel: the stage-0 vocoder convs and
// every mid-size acoustic GEMM are L2->SM-bandwidth-bound with 128 x 256 tiles (each CTA pulls 16 KB of
// activations + 32 KB of weights per K step, ~57 B/clk/SM).  Here a cluster of two CTAs (a TPC pair)
// computes a 256 x 256 tile: each CTA loads ITS 128 rows of activations and HALF of the weight tile
// (128 of the 256 output channels); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads
// both halves from the two CTAs' shared memory and writes each CTA's 128 rows into that CTA's TMEM.
// 32 KB per CTA per K step instead of 48 KB.  Barriers: TMA loads of both CTAs complete on the LEADER's
// full barrier; the leader's commits multicast to both CTAs' empty / tfull barriers; both CTAs'
// epilogue warps arrive on the leader's tempty barrier.
// ---------------------------------------------------------------------------------------------
template <int BK, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CV_THREADS, 1)
conv_igemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWh,
                       const __grid_constant__ EpiMaps emaps, const __grid_constant__ ConvArgs a) {
  constexpr int BN = 256;
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr int WH_BYTES = (BN / 2) * BK * 2;          // this CTA's half of the weight tile
  constexpr int STAGE_BYTES = A_BYTES + WH_BYTES;
  constexpr int TMEM_COLS = 512;

  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = ((smem_u32(smem_dyn) + 1023u) & ~1023u) + (uint32_t)a.pad_stages * STAGE_BYTES;   // pad: experiments
  const int S = a.stages;
  const uint32_t bar_base = smem_base + S * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * S + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * S + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bar_base + 8u * (2 * S + 6) - smem_u32(smem_dyn)));
  for (int i = threadIdx.x; i < a.CoutP; i += blockDim.x) bias_s[i] = (a.bias != nullptr && i < a.Cout) ? a.bias[i] : 0.f;
  const uint32_t epi_base = (bar_base + 8u * (2 * S + 6) + (uint32_t)a.CoutP * 4u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = a.B * a.n_ttiles * a.n_ftiles;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_super = m_pairs * (a.CoutP / BN);
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int k_iters = a.ntaps * a.kchunks;

  if (threadIdx.x == 0) {
    // full: one arrive (the leader's expect_tx) + the bytes of both CTAs; empty / tfull: the leader's commits;
    // tempty (used in the leader only): 4 epilogue warps of each CTA
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); }
    if (a.epi_tma) for (int i = 0; i < 8; ++i) mbar_init(epi_base + 8u * epi_warp_bytes(BN) + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWh) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();            // both CTAs' barriers are initialised and their TMEM is allocated (also a CTA barrier)
  tc_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own activation rows + own half of the weight tile =====
      int it = 0;
      for (int st = cid; st < total_super; st += n_clusters) {
        if (conv_unit_skip(a, st % m_pairs, 2)) continue;
        int mt = (st % m_pairs) * 2 + (int)rank;
        const int n0 = (st / m_pairs) * BN;
        const int ft = mt % a.n_ftiles; mt /= a.n_ftiles;
        const int tt = mt % a.n_ttiles; mt /= a.n_ttiles;
        const int b = mt, t0 = tt * a.tT, f0 = ft * a.tF;          // b == B for the padding tile: zero-filled by TMA
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (leader) mbar_expect_tx(full_bar(s), 2u * STAGE_BYTES);
          const uint32_t fb = mapa_u32(full_bar(s), 0);            // the leader's full barrier
          const int tap = kit / a.kchunks, kc = kit - tap * a.kchunks;
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          tma2_load_4d(sa, &tmA, fb, kc * BK, f0 + a.tap_df[tap], t0 + a.tap_dt[tap], b);
          tma2_load_2d(sa + A_BYTES, &tmWh, fb, kc * BK, tap * a.CoutP + n0 + (int)rank * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===== MMA issuer (leader CTA only) =====
      int it = 0, lt = 0;
      for (int st = cid; st < total_super; st += n_clusters) {
        if (conv_unit_skip(a, st % m_pairs, 2)) continue;
        const int acc = lt & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint64_t da = make_smem_desc<BK>(sa);
          const uint64_t db = make_smem_desc<BK>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc2_mma_f16(tacc, da + uint64_t(2 * k), db + uint64_t(2 * k), a.idesc, (kit | k) != 0 ? 1u : 0u);
          tc2_commit(empty_bar(s));
        }
        tc2_commit(tfull_bar(acc));
        ++lt;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs, each on its own 128 rows): arrives on the leader's tempty barriers =====
    const uint32_t tempty_leader = mapa_u32(tempty_bar(0), 0);
    if (a.epi_tma == 2)
      run_epilogue_tma<BN, BF16, true>(a, emaps, tmem_base, tfull_bar(0), tempty_leader, warp, lane, m_pairs, total_super, bias_s,
                                       epi_base, cid, n_clusters, 2, (int)rank, true);
    else if (a.epi_tma)
      run_epilogue_tma<BN, BF16>(a, emaps, tmem_base, tfull_bar(0), tempty_leader, warp, lane, m_pairs, total_super, bias_s, epi_base,
                                 cid, n_clusters, 2, (int)rank, true);
    else
      run_epilogue<BN, BF16>(a, tmem_base, tfull_bar(0), tempty_leader, warp, lane, m_pairs, total_super, bias_s,
                             cid, n_clusters, 2, (int)rank, true);
  }

  tc_fence_before();
  cluster_sync();            // the peer may still multicast commits into this CTA / read its operands
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow-channel 1-D variant (Cin, Cout <= 64: vocoder stages with 64 / 32 channels, 37 % of its
// FLOPs but > 50 % of its time with the kernel above, which re-loads the activation tile once per
// tap).  Here the activation tile is loaded ONCE per output tile with its halo
// (128 + (k-1)*dil rows) and ALL taps' weights stay resident in shared memory for the CTA's life:
// L2->smem traffic per tile drops from k*(A+W) to ~1.4*A.
//
// Layout: un-swizzled K-major "core matrix" planes.  One TMA box {8 ch, HRP rows, Cin/8 planes}
// of a tensor map whose dimensions are ordered (8 channels, time, channel-block) lands in shared
// memory as [plane][row][8 ch]: every 16-byte unit is one row of an 8x8 core matrix, 8 rows are
// contiguous (128 B), so a tap is simply a start-address offset of `rows * 16` bytes in the UMMA
// descriptor (LBO = plane stride, SBO = 128 B) — no swizzle phase to keep aligned.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes) {
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}

template <int BN, bool BF16>
__global__ void __launch_bounds__(CV_THREADS, 2)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ EpiMaps emaps,
                 const __grid_constant__ ConvArgs a) {
  constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 127u) & ~127u;
  const int S = a.stages;
  const int KC = a.kchunks;                 // 8-channel planes
  const int HRP = a.halo_rows;              // rows per plane of the A tile (multiple of 8)
  const uint32_t a_bytes = (uint32_t)KC * HRP * 16;
  const uint32_t w_tap_bytes = (uint32_t)KC * BN * 16;
  const uint32_t w_base = smem_base + S * a_bytes;
  const uint32_t bar_base = (w_base + a.ntaps * w_tap_bytes + 7u) & ~7u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * S + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * S + 2 + i); };
  const uint32_t w_bar = bar_base + 8u * (2 * S + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 5);
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bar_base + 8u * (2 * S + 6) - smem_u32(smem_dyn)));
  for (int i = threadIdx.x; i < a.CoutP; i += blockDim.x) bias_s[i] = (a.bias != nullptr && i < a.Cout) ? a.bias[i] : 0.f;
  const uint32_t epi_base = (bar_base + 8u * (2 * S + 6) + (uint32_t)a.CoutP * 4u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = a.B * a.n_ttiles;
  const int total_tiles = m_tiles;          // single N tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    mbar_init(w_bar, 1);
    if (a.epi_tma) for (int i = 0; i < 8; ++i) mbar_init(epi_base + 8u * epi_warp_bytes(BN) + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // barrier init / TMEM alloc / descriptor prefetch above overlap the previous kernel's tail
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: weights once, then one halo tile per output tile =====
      mbar_expect_tx(w_bar, a.ntaps * w_tap_bytes);
      for (int tap = 0; tap < a.ntaps; ++tap)
        tma_load_4d(w_base + tap * w_tap_bytes, &tmW, w_bar, 0, tap * a.CoutP, 0, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int tt = tile % a.n_ttiles, b = tile / a.n_ttiles;
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), a_bytes);
        tma_load_4d(smem_base + s * a_bytes, &tmA, full_bar(s), 0, tt * 128 - a.pad_lo, 0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elected lane of a converged warp: back-to-back UTCHMMA issue (no per-instruction election loop)
      // ===== MMA issuer =====
      mbar_wait(w_bar, 0);
      int lt = 0;
      const int ksteps = KC / 2;             // 16 channels per MMA
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const int s = lt % S;
        const uint32_t ph = (lt / S) & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)lt >> 1) & 1u) ^ 1u);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        const uint32_t sa = smem_base + s * a_bytes;
        uint32_t first = 0;
        for (int tap = 0; tap < a.ntaps; ++tap) {
          const uint32_t arow = sa + (uint32_t)(a.pad_lo + a.tap_dt[tap]) * 16u;
          const uint32_t wtap = w_base + tap * w_tap_bytes;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t da = make_desc_noswz(arow + (uint32_t)(2 * j * HRP) * 16u, (uint32_t)HRP * 16u);
            const uint64_t db = make_desc_noswz(wtap + (uint32_t)(2 * j * BN) * 16u, (uint32_t)BN * 16u);
            tc_mma_f16(tacc, da, db, a.idesc, first);
            first = 1u;
          }
        }
        tc_commit(empty_bar(s));
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    if (a.epi_tma)
      run_epilogue_tma<BN, BF16>(a, emaps, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s, epi_base);
    else
      run_epilogue<BN, BF16>(a, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Swizzled halo variant for 64 / 128 input channels (vocoder stages 1-2).  ncu showed the per-tap
// kernel is bound by L2->SM bandwidth there (lts ~52 %, 44 B/clk/SM: the LTS cap): each tap re-loads
// the activation tile and every tile re-streams the weights.  Here the activation tile (with halo)
// is loaded once per output tile as 64-channel chunks in the standard 128B-swizzled K-major layout
// (full-width 128 B TMA rows) and all taps' weights are resident in shared memory.  A tap is a
// start-address offset of r*128 B into the swizzled tile.  Measured on B200: the UMMA swizzle is a
// function of the absolute shared-memory address bits (like TMA's), so an unaligned start row
// needs NO base-offset in the descriptor (setting bits 49..51 to r & 7 gives wrong results;
// tests/test_kernels_gpu.py covers row offsets 1..25).
// ---------------------------------------------------------------------------------------------
template <int BN, int BKC, bool BF16>
__global__ void __launch_bounds__(CV_THREADS, 2)
conv_halo_sw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ EpiMaps emaps,
                    const __grid_constant__ ConvArgs a) {
  constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int S = a.stages;
  constexpr uint32_t RB = BKC * 2u;          // bytes per row of a chunk: 128 (128B swizzle) or 64 (64B swizzle)
  const int KCH = a.kchunks;                // BKC-channel chunks
  const int HRP = a.halo_rows;              // rows of the halo tile (multiple of 8)
  const uint32_t chunk_bytes = (uint32_t)HRP * RB;
  const uint32_t a_bytes = (uint32_t)KCH * chunk_bytes;
  const uint32_t w_blk = (uint32_t)BN * RB;                         // one (tap, chunk) weight block
  const uint32_t w_base = smem_base + S * a_bytes;
  const uint32_t bar_base = w_base + a.ntaps * KCH * w_blk;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * S + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * S + 2 + i); };
  const uint32_t w_bar = bar_base + 8u * (2 * S + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 5);
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bar_base + 8u * (2 * S + 6) - smem_u32(smem_dyn)));
  for (int i = threadIdx.x; i < a.CoutP; i += blockDim.x) bias_s[i] = (a.bias != nullptr && i < a.Cout) ? a.bias[i] : 0.f;
  const uint32_t epi_base = (bar_base + 8u * (2 * S + 6) + (uint32_t)a.CoutP * 4u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = a.B * a.n_ttiles;
  const int total_tiles = m_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    mbar_init(w_bar, 1);
    if (a.epi_tma) for (int i = 0; i < 8; ++i) mbar_init(epi_base + 8u * epi_warp_bytes(BN) + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // barrier init / TMEM alloc / descriptor prefetch above overlap the previous kernel's tail
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, a.ntaps * KCH * w_blk);
      for (int tap = 0; tap < a.ntaps; ++tap)
        for (int c = 0; c < KCH; ++c)
          tma_load_2d(w_base + (tap * KCH + c) * w_blk, &tmW, w_bar, c * BKC, tap * a.CoutP);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int tt = tile % a.n_ttiles, b = tile / a.n_ttiles;
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), a_bytes);
        for (int c = 0; c < KCH; ++c)
          tma_load_4d(smem_base + s * a_bytes + c * chunk_bytes, &tmA, full_bar(s), c * BKC, 0, tt * 128 - a.pad_lo, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elected lane of a converged warp: back-to-back UTCHMMA issue (no per-instruction election loop)
      mbar_wait(w_bar, 0);
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const int s = lt % S;
        const uint32_t ph = (lt / S) & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)lt >> 1) & 1u) ^ 1u);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        const uint32_t sa = smem_base + s * a_bytes;
        uint32_t first = 0;
        for (int tap = 0; tap < a.ntaps; ++tap) {
          const uint32_t r = (uint32_t)(a.pad_lo + a.tap_dt[tap]);
          const uint64_t boff = a.halo_baseoff ? (uint64_t(r & 7u) << 49) : 0ull;
          for (int c = 0; c < KCH; ++c) {
            const uint64_t da = make_smem_desc<BKC>(sa + c * chunk_bytes + r * RB) | boff;
            const uint64_t db = make_smem_desc<BKC>(w_base + (tap * KCH + c) * w_blk);
#pragma unroll
            for (int k = 0; k < BKC / 16; ++k) {
              tc_mma_f16(tacc, da + uint64_t(2 * k), db + uint64_t(2 * k), a.idesc, first);
              first = 1u;
            }
          }
        }
        tc_commit(empty_bar(s));
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    if (a.epi_tma)
      run_epilogue_tma<BN, BF16>(a, emaps, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s, epi_base);
    else
      run_epilogue<BN, BF16>(a, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, m_tiles, total_tiles, bias_s);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

inline int pick_tile_n(int Cout) {
  if (Cout <= 16) return 16;
  if (Cout <= 32) return 32;
  if (Cout <= 64) return 64;
  if (Cout <= 128) return 128;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  // minimise padding between 128- and 256-wide tiles, prefer the wider one on ties
  int p256 = (Cout + 255) / 256 * 256, p128 = (Cout + 127) / 128 * 128;
  return p128 < p256 ? 128 : 256;
}


template <int BN, int BK, bool BF16>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmW, const EpiMaps& em, ConvArgs& a,
                       int total_tiles, cudaStream_t st) {
  constexpr int STAGE_BYTES = 128 * BK * 2 + BN * BK * 2;
  // BN = 256 needs all 512 TMEM columns (two accumulator stages): one CTA per SM with a deep ring.
  // Narrower tiles run two CTAs per SM (TMEM 2*BN <= 256 columns each, ~100 KB of ring each).
  // The TMA epilogue needs 64 KB of staging per CTA: one CTA per SM with a deep ring instead of two.
  const int ctas_per_sm = (BN >= 256 || a.epi_tma) ? 1 : 2;
  const int budget = a.epi_tma ? (222 * 1024 - (int)epi_bytes(BN) - 2048 - a.CoutP * 4)
                               : ((BN >= 256 ? 208 : 104) * 1024 - a.CoutP * 4);   // ring + bias within 227 KB / SM
  int stages = budget / STAGE_BYTES;
  if (stages > 10) stages = 10;
  if (stages < 2) stages = 2;
  a.stages = stages;
  static const bool no_skip = getenv("ASB_CONV_NO_SKIP") != nullptr;
  a.skip = (a.lens != nullptr && !no_skip) ? 1 : 0;     // ragged batches: tiles beyond an item's length are left out
  const size_t smem = (size_t)stages * STAGE_BYTES + 8 * (2 * stages + 6) + (size_t)a.CoutP * 4 + 1024 + 16 +
                      (a.epi_tma ? epi_bytes(BN) + 1024 : 0);
  ASB_SMEM_OPT_IN(227 * 1024, conv_igemm_kernel<BN, BK, BF16>);
  int grid = num_sms() * ctas_per_sm;
  if (grid > total_tiles) grid = total_tiles;
  ASB_CUDA(launch_k(conv_igemm_kernel<BN, BK, BF16>, grid, CV_THREADS, smem, st, tmA, tmW, em, a));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

template <int BK, bool BF16>
int launch_conv_2cta(const CUtensorMap& tmA, const CUtensorMap& tmWh, const EpiMaps& em, ConvArgs& a, cudaStream_t st) {
  constexpr int BN = 256;
  constexpr int STAGE_BYTES = 128 * BK * 2 + (BN / 2) * BK * 2;
  const int budget = a.epi_tma ? (222 * 1024 - (int)epi_bytes(BN) - 2048 - a.CoutP * 4) : (208 * 1024 - a.CoutP * 4);
  int stages = budget / STAGE_BYTES;
  // At most 4 ring stages.  With 6 (what the budget gives launches with the direct epilogue) the kernel raised a rare
  // sticky "out-of-range shared address" fault in the leader's MMA issuer whenever its ring ran completely full --
  // epilogue-bound launches (fp32 residual + two outputs) overlapped with other work; ~1 in 6000 launches of the serving
  // step, within 500 launches of tools/stress_two2cta.py.  On identical kernel code 6 stages failed 5 / 5 runs and 2 - 5
  // stages passed 13 / 13 (30 000 launches each of the repro, 4000 serving steps); every launch with the TMA epilogue
  // already ran 4 stages and none was ever implicated (flight recorder + compute-sanitizer).  Root
  // cause not established: the fault is imprecise (reported on the issuer while it spins on a full barrier), which points
  // at an asynchronous operation it issued -- with a full 6-stage ring up to 7 multicast tcgen05.commit arrivals
  // (6 ring slots + the accumulator barrier) are outstanding at once.  ASB_2CTA_STAGES overrides the cap (repro only).
  static const int max_stages = getenv("ASB_2CTA_STAGES") ? atoi(getenv("ASB_2CTA_STAGES")) : 4;
  if (stages > max_stages) stages = max_stages;
  if (stages < 2) stages = 2;
  a.stages = stages;
  static const int pad = getenv("ASB_2CTA_PAD") ? atoi(getenv("ASB_2CTA_PAD")) : 0;
  a.pad_stages = (stages + pad) * STAGE_BYTES <= budget ? pad : 0;
  a.idesc = (a.idesc & ~(0x1Fu << 24)) | (uint32_t(256 >> 4) << 24);      // UMMA M = 256 across the CTA pair
  static const bool no_skip = getenv("ASB_CONV_NO_SKIP") != nullptr;
  a.skip = (a.lens != nullptr && !no_skip) ? 1 : 0;
  const size_t smem = (size_t)(stages + a.pad_stages) * STAGE_BYTES + 8 * (2 * stages + 6) + (size_t)a.CoutP * 4 + 1024 + 16 +
                      (a.epi_tma ? epi_bytes(BN) + 1024 : 0);
  ASB_SMEM_OPT_IN(227 * 1024, conv_igemm_2cta_kernel<BK, BF16>);
  const int m_tiles = a.B * a.n_ttiles * a.n_ftiles;
  const int total_super = ((m_tiles + 1) / 2) * (a.CoutP / BN);
  int clusters = num_sms() / 2;
  if (clusters > total_super) clusters = total_super;
  ASB_CUDA(launch_k(conv_igemm_2cta_kernel<BK, BF16>, 2 * clusters, CV_THREADS, smem, st, tmA, tmWh, em, a));
  return AS_OK;
}

template <int BN, bool BF16>
int launch_halo(const as_conv_params* p, ConvArgs& a, const EpiMaps& em, EncodeTiledFn enc, cudaStream_t st) {
  const int KC = p->Cin / 8;
  int lo = 0, hi = 0;
  for (int j = 0; j < p->ntaps; ++j) { lo = p->tap_dt[j] < lo ? p->tap_dt[j] : lo; hi = p->tap_dt[j] > hi ? p->tap_dt[j] : hi; }
  const int HRP = (128 + hi - lo + 7) / 8 * 8;
  a.kchunks = KC; a.halo_rows = HRP; a.pad_lo = -lo; a.skip = 0;
  a.tT = 128; a.tF = 1; a.n_ttiles = (p->To + 127) / 128; a.n_ftiles = 1;
  const size_t a_bytes = (size_t)KC * HRP * 16, w_bytes = (size_t)p->ntaps * KC * BN * 16;
  int stages = 4;
  while (stages > 2 && stages * a_bytes + w_bytes > 190 * 1024) --stages;
  a.stages = stages;
  const size_t epi = a.epi_tma ? epi_bytes(BN) + 1024 : 0;
  while (stages > 2 && stages * a_bytes + w_bytes + epi > 208 * 1024) --stages;
  a.stages = stages;
  const size_t smem = stages * a_bytes + w_bytes + 8 * (2 * stages + 6) + (size_t)a.CoutP * 4 + 256 + 16 + epi;
  const CUtensorMapDataType dt = BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmW;
  {
    // dims ordered (8 channels, time, channel block, batch): smem gets [plane][row][8ch]
    cuuint64_t dims[4] = {8, (cuuint64_t)p->T, (cuuint64_t)KC, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, 16, (cuuint64_t)p->x_ld * 2 * p->T};
    cuuint32_t box[4] = {8, (cuuint32_t)HRP, (cuuint32_t)KC, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, dt, 4, const_cast<void*>(p->x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(halo A) failed: %d", (int)r);
  }
  {
    // packed weights [ntaps*CoutP rows][CinP]: (8 ci, rows, ci block, 1) -> [plane][co][8ci] per tap
    cuuint64_t dims[4] = {8, (cuuint64_t)p->ntaps * p->CoutP, (cuuint64_t)KC, 1};
    cuuint64_t strides[3] = {(cuuint64_t)p->CinP * 2, 16, (cuuint64_t)p->CinP * 2 * p->ntaps * p->CoutP};
    cuuint32_t box[4] = {8, (cuuint32_t)BN, (cuuint32_t)KC, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmW, dt, 4, const_cast<void*>(p->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(halo W) failed: %d", (int)r);
  }
  ASB_SMEM_OPT_IN(227 * 1024, conv_halo_kernel<BN, BF16>);
  const int total_tiles = p->B * a.n_ttiles;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > total_tiles) grid = total_tiles;
  ASB_CUDA(launch_k(conv_halo_kernel<BN, BF16>, grid, CV_THREADS, smem, st, tmA, tmW, em, a));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

template <int BN, int BKC, bool BF16>
int launch_halo_sw(const as_conv_params* p, ConvArgs& a, const EpiMaps& em, EncodeTiledFn enc, cudaStream_t st) {
  const int KCH = p->Cin / BKC;
  int lo = 0, hi = 0;
  for (int j = 0; j < p->ntaps; ++j) { lo = p->tap_dt[j] < lo ? p->tap_dt[j] : lo; hi = p->tap_dt[j] > hi ? p->tap_dt[j] : hi; }
  const int HRP = (128 + hi - lo + 7) / 8 * 8;
  a.kchunks = KCH; a.halo_rows = HRP; a.pad_lo = -lo; a.skip = 0;
  static const int baseoff_mode = getenv("ASB_HALO_BASEOFF") ? atoi(getenv("ASB_HALO_BASEOFF")) : 0;
  a.halo_baseoff = baseoff_mode;
  a.tT = 128; a.tF = 1; a.n_ttiles = (p->To + 127) / 128; a.n_ftiles = 1;
  const size_t a_bytes = (size_t)KCH * HRP * BKC * 2, w_bytes = (size_t)p->ntaps * KCH * BN * BKC * 2;
  int stages = 4;
  while (stages > 2 && stages * a_bytes + w_bytes > 200 * 1024) --stages;
  a.stages = stages;
  const size_t epi = a.epi_tma ? epi_bytes(BN) + 1024 : 0;
  while (stages > 2 && stages * a_bytes + w_bytes + epi > 208 * 1024) --stages;
  // two CTAs per SM (more tiles in flight) when the resident weights leave room for it
  if (2 * a_bytes + w_bytes + epi + 4096 <= 110 * 1024) { stages = (int)((110 * 1024 - w_bytes - epi - 4096) / a_bytes); if (stages > 4) stages = 4; }
  a.stages = stages;
  const size_t smem = stages * a_bytes + w_bytes + 8 * (2 * stages + 6) + (size_t)a.CoutP * 4 + 1024 + 16 + epi;
  const CUtensorMapDataType dt = BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->Cin, 1, (cuuint64_t)p->T, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, (cuuint64_t)p->x_ld * 2, (cuuint64_t)p->x_ld * 2 * p->T};
    cuuint32_t box[4] = {BKC, 1, (cuuint32_t)HRP, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, dt, 4, const_cast<void*>(p->x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     BKC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(halo-sw A) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p->CinP, (cuuint64_t)p->ntaps * p->CoutP};
    cuuint64_t strides[1] = {(cuuint64_t)p->CinP * 2};
    cuuint32_t box[2] = {BKC, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmW, dt, 2, const_cast<void*>(p->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     BKC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASB_REQUIRE(r == CUDA_SUCCESS, AS_ERR_CUDA, "cuTensorMapEncodeTiled(halo-sw W) failed: %d", (int)r);
  }
  ASB_SMEM_OPT_IN(227 * 1024, conv_halo_sw_kernel<BN, BKC, BF16>);
  const int total_tiles = p->B * a.n_ttiles;
  int per_sm = (int)((226 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > total_tiles) grid = total_tiles;
  ASB_CUDA(launch_k(conv_halo_sw_kernel<BN, BKC, BF16>, grid, CV_THREADS, smem, st, tmA, tmW, em, a));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

}  // namespace asb
