// Memory-bound kernels of the synthesis path: embedding, LayerNorm, InstanceNorm statistics,
// AdaIN(+LeakyReLU, + depthwise ConvTranspose "pool"), nearest upsample, length regulation,
// small-Cin direct convolution, depthwise convolution, pooling, layout changes.
// All of them are coalesced along the channel axis of the channels-last activations and
// vectorised where the row pitch allows it; reductions use warp shuffles.
#include "tc_util.cuh"

namespace asb {

// Grid-stride loop over `total` elements with 32-bit index arithmetic whenever the count allows it:
// 64-bit div/mod is emulated (~100 instructions) and dominated these kernels (ncu: conv_small 530 us).
#define ASB_GRID_STRIDE(i, total, ...)                                                              \
  if ((total) <= 0x7fffffffLL) {                                                                    \
    const unsigned _n = (unsigned)(total);                                                          \
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < _n; i += gridDim.x * blockDim.x)   \
      __VA_ARGS__                                                                                   \
  } else {                                                                                          \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total);               \
         i += (long long)gridDim.x * blockDim.x)                                                    \
      __VA_ARGS__                                                                                   \
  }


static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// embedding * scale  (Utils/RelTransformerEnc.py:372-373)
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table,
                             int n_vocab, int B, int T, int C, float scale,
                             const int* __restrict__ lens, float* out32, void* out16, int dt16) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row % T;
  const bool valid = lens == nullptr || t < lens[b];
  long long id = tok[row];
  if (id < 0 || id >= n_vocab) id = 0;
  const float* src = table + id * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = valid ? src[c] * scale : 0.f;
    if (out32) out32[row * C + c] = v;
    if (out16) reinterpret_cast<uint16_t*>(out16)[row * C + c] = to16(v, dt16);
  }
}

// ---------------------------------------------------------------------------------------------
// channel LayerNorm, one warp per row (C <= 1024)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAX_PER_LANE = 32;

__global__ void layernorm_kernel(const void* __restrict__ x, int xdt, long long x_ld, long long rows,
                                 int T, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int act, float slope,
                                 const int* __restrict__ lens, void* oa, int oadt, long long oa_ld,
                                 void* ob, int obdt, long long ob_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row % T;
  const bool valid = lens == nullptr || t < lens[b];
  float v[LN_MAX_PER_LANE];
  float s = 0.f;
  const int n = (C + 31) / 32;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    v[i] = 0.f;
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) { v[i] = ldany(x, row * x_ld + c, xdt); s += v[i]; }
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) { float d = v[i] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) {
        float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
        y = valid ? apply_act(y, act, slope) : 0.f;
        if (oa) stany(oa, row * oa_ld + c, y, oadt);
        if (ob) stany(ob, row * ob_ld + c, y, obdt);
      }
    }
  }
}

__device__ __forceinline__ void st4any(void* p, long long off, float4 v, int dt);   // defined below

// fp32 rows with C a multiple of 128 (the encoders' 512 / 256 / 1024 channels): each lane owns C / 128 groups of 4
// consecutive channels -- 16-byte loads, all in flight before the first use, and 8- / 16-byte stores (the scalar
// kernel above issues 16 four-byte loads and up to 32 stores per lane at C = 512)
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_vec4_kernel(const float* __restrict__ x, long long x_ld, long long rows, int T, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, int act, float slope, const int* __restrict__ lens,
                      void* oa, int oadt, long long oa_ld, void* ob, int obdt, long long ob_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row % T;
  const bool valid = lens == nullptr || t < lens[b];
  float4 v[NV];
  const float4* xr = reinterpret_cast<const float4*>(x + row * x_ld) + lane;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldg(xr + 32 * i);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 4 * lane + 128 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + be.x; y.y = (v[i].y - mean) * rstd * g.y + be.y;
    y.z = (v[i].z - mean) * rstd * g.z + be.z; y.w = (v[i].w - mean) * rstd * g.w + be.w;
    if (valid) { y.x = apply_act(y.x, act, slope); y.y = apply_act(y.y, act, slope); y.z = apply_act(y.z, act, slope); y.w = apply_act(y.w, act, slope); }
    else y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oa) st4any(oa, row * oa_ld + c, y, oadt);
    if (ob) st4any(ob, row * ob_ld + c, y, obdt);
  }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm statistics: block = (b, 32 channels), 32 x 8 threads, two passes (mean, then
// centred second moment) like the reference's biased variance.
// ---------------------------------------------------------------------------------------------
__global__ void instnorm_stats_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T,
                                      int C, const int* __restrict__ lens, float eps,
                                      float* __restrict__ stats) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const int len = lens ? min(lens[b], T) : T;
  const long long base = (long long)b * T * x_ld;
  float s = 0.f;
  if (c < C)
    for (int t = ty; t < len; t += 8) s += ldany(x, base + (long long)t * x_ld + c, xdt);
  red[ty][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
  const float mean = len > 0 ? tot / len : 0.f;
  __syncthreads();
  float q = 0.f;
  if (c < C)
    for (int t = ty; t < len; t += 8) {
      float d = ldany(x, base + (long long)t * x_ld + c, xdt) - mean;
      q += d * d;
    }
  red[ty][threadIdx.x] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += red[i][threadIdx.x];
    var = len > 0 ? var / len : 0.f;
    stats[((long long)b * C + c) * 2 + 0] = mean;
    stats[((long long)b * C + c) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------------------------------------
// AdaIN + LeakyReLU (+ depthwise ConvTranspose1d k3 s2 p1 op1)
// ---------------------------------------------------------------------------------------------
__global__ void adain_apply_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                   int C, const float* __restrict__ stats,
                                   const float* __restrict__ gb, long long gb_ld, float slope,
                                   const int* __restrict__ lens, const float* __restrict__ up_w,
                                   const float* __restrict__ up_b, void* out, int odt,
                                   long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const int To = up_w ? 2 * T : T;
  const long long total = (long long)B * To * C;
  ASB_GRID_STRIDE(i, total, {
    const int c = i % C;
    const auto r = i / C;
    const int to = r % To;
    const int b = r / To;
    const int len = lens ? min(lens[b], T) : T;
    const float mean = stats[((long long)b * C + c) * 2], rstd = stats[((long long)b * C + c) * 2 + 1];
    const float g = 1.f + gb[(long long)b * gb_ld + c], be = gb[(long long)b * gb_ld + C + c];
    auto a_at = [&](int t) -> float {
      if (t >= len) return 0.f;
      float v = ldany(x, ((long long)b * T + t) * x_ld + c, xdt);
      v = (v - mean) * rstd * g + be;
      return v > 0.f ? v : v * slope;
    };
    float y;
    if (up_w == nullptr) {
      y = a_at(to);
    } else {
      const int m = to >> 1;
      if (m >= len) y = 0.f;
      else if ((to & 1) == 0) y = a_at(m) * up_w[c * 3 + 1] + up_b[c];
      else y = a_at(m) * up_w[c * 3 + 2] + a_at(m + 1) * up_w[c * 3 + 0] + up_b[c];
    }
    stany(out, ((long long)b * To + to) * out_ld + c, y, odt);
  })
}

// Vector form of the kernel above (4 channels per lane, per-channel constants in registers, a warp
// reads / writes one contiguous row segment): the norm + affine + LeakyReLU pass is pure HBM streaming.
// grid = (C / 128, row chunks, B); requires C % 4 == 0 and 16- / 8-byte aligned rows.
constexpr int AD_ROWS = 32;   // input rows per CTA (8 warps x 4 rows)

__device__ __forceinline__ float4 ld4any(const void* p, long long off, int dt) {
  if (dt == AS_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + off));
  return make_float4(from16(uint16_t(u.x & 0xFFFF), dt), from16(uint16_t(u.x >> 16), dt),
                     from16(uint16_t(u.y & 0xFFFF), dt), from16(uint16_t(u.y >> 16), dt));
}
__device__ __forceinline__ void st4any(void* p, long long off, float4 v, int dt) {
  if (dt == AS_F32) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + off) = v; return; }
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + off) = make_uint2(pack16(v.x, v.y, dt), pack16(v.z, v.w, dt));
}

template <bool UP>
__global__ void __launch_bounds__(256)
adain_apply_vec_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T, int C,
                       const float* __restrict__ stats, const float* __restrict__ gb, long long gb_ld,
                       float slope, const int* __restrict__ lens, const float* __restrict__ up_w,
                       const float* __restrict__ up_b, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int c = (blockIdx.x * 32 + lane) * 4;
  if (c >= C) return;
  const int len = lens ? min(lens[b], T) : T;
  float sc[4], mu[4], sh[4], w0[4], w1[4], w2[4], ub[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float mean = stats[((long long)b * C + c + e) * 2], rstd = stats[((long long)b * C + c + e) * 2 + 1];
    const float g = 1.f + gb[(long long)b * gb_ld + c + e], be = gb[(long long)b * gb_ld + C + c + e];
    sc[e] = rstd * g;
    mu[e] = mean;           // subtract the mean first: bias-dominated channels (|mean| >> std) would cancel badly otherwise
    sh[e] = be;
    if (UP) { w0[e] = up_w[(c + e) * 3]; w1[e] = up_w[(c + e) * 3 + 1]; w2[e] = up_w[(c + e) * 3 + 2]; ub[e] = up_b[c + e]; }
  }
  auto act_row = [&](int t) -> float4 {
    if (t >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = ld4any(x, ((long long)b * T + t) * x_ld + c, xdt);
    v.x = (v.x - mu[0]) * sc[0] + sh[0]; v.y = (v.y - mu[1]) * sc[1] + sh[1];
    v.z = (v.z - mu[2]) * sc[2] + sh[2]; v.w = (v.w - mu[3]) * sc[3] + sh[3];
    v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
    return v;
  };
  const int t_end = min(T, (int)(blockIdx.y + 1) * AD_ROWS);
  for (int t = blockIdx.y * AD_ROWS + warp; t < t_end; t += 8) {
    const float4 a0 = act_row(t);
    if (!UP) {
      st4any(out, ((long long)b * T + t) * out_ld + c, a0, odt);
    } else {
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), od = ev;
      if (t < len) {
        const float4 a1 = act_row(t + 1);
        ev = make_float4(a0.x * w1[0] + ub[0], a0.y * w1[1] + ub[1], a0.z * w1[2] + ub[2], a0.w * w1[3] + ub[3]);
        od = make_float4(a0.x * w2[0] + a1.x * w0[0] + ub[0], a0.y * w2[1] + a1.y * w0[1] + ub[1],
                         a0.z * w2[2] + a1.z * w0[2] + ub[2], a0.w * w2[3] + a1.w * w0[3] + ub[3]);
      }
      st4any(out, ((long long)b * 2 * T + 2 * t) * out_ld + c, ev, odt);
      st4any(out, ((long long)b * 2 * T + 2 * t + 1) * out_ld + c, od, odt);
    }
  }
}

// InstanceNorm statistics + AdaIN + LeakyReLU (+ pool) in ONE pass over HBM: a CTA owns 32 channels of one
// item, keeps the whole [T, 32] fp32 slab in shared memory (T <= 1536), computes the exact two-pass
// mean / biased variance from it and writes the activated tensor.  Replaces the stats + apply launches
// (3 reads + 1 write of the tensor -> 1 read + 1 write).
constexpr int ADF_CS = 4;                 // CTAs per cluster: the T axis of one (item, 32-channel block) is split 4 ways
constexpr int ADF_MAX_TR = 1536;          // rows per CTA that fit in shared memory -> T <= 6144

// 4 consecutive channels of one row: 16-byte (fp32) or 8-byte (16-bit) load, dtype resolved at compile time
template <int XDT>
__device__ __forceinline__ float4 adf_load4(const void* x, long long off) {
  if (XDT == AS_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + off));
  return make_float4(from16(uint16_t(u.x & 0xFFFF), XDT), from16(uint16_t(u.x >> 16), XDT),
                     from16(uint16_t(u.y & 0xFFFF), XDT), from16(uint16_t(u.y >> 16), XDT));
}
__device__ __forceinline__ uint32_t adf_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float adf_ld_peer(const float* p, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void adf_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One thread-block CLUSTER of 4 CTAs per (item, 32-channel block); CTA r owns rows [r*TR, (r+1)*TR) and keeps
// them in shared memory.  Exact two-pass statistics: per-CTA partial sums, exchanged through distributed
// shared memory (ld.shared::cluster) around two cluster barriers; then each CTA normalises and writes its
// rows.  Small slabs (28 KB at T = 800) put 8 CTAs on an SM, so loads, reductions and stores of different
// CTAs overlap and the pass streams HBM (one read + one write of the tensor).
// Thread layout: warp w, lane -> (row-in-group r4 = lane >> 3, channel quad c4 = lane & 7): one warp
// instruction moves 4 rows x 128 B; ADF_U independent 16-byte loads per thread are in flight.
template <bool UP, int XDT>
__global__ void __cluster_dims__(ADF_CS, 1, 1) __launch_bounds__(256)
adain_fused_kernel(const void* __restrict__ x, long long x_ld, int T, int TR, int C,
                   const float* __restrict__ gb, long long gb_ld, float eps, float slope,
                   const int* __restrict__ lens, const float* __restrict__ up_w, const float* __restrict__ up_b,
                   void* out, int odt, long long out_ld, float* __restrict__ stats_out) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(16) float slab[];            // [TR][32]
  __shared__ float red[8][32];
  __shared__ float part[2][32];                            // this CTA's partial sum / partial squared deviation
  constexpr int ADF_U = 8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r4 = lane >> 3, c4 = lane & 7;
  const uint32_t rank = adf_cluster_rank();
  const int b = blockIdx.y, c = (blockIdx.x / ADF_CS) * 32 + 4 * c4;
  const bool cok = c < C;                                   // C % 4 == 0: a quad is all-in or all-out
  const int len = lens ? min(lens[b], T) : T;
  const int t_lo = (int)rank * TR, t_hi = min(len, t_lo + TR);   // valid rows of this CTA
  const long long base = (long long)b * T * x_ld + c;
  const int row0 = t_lo + w * 4 + r4;                       // this thread's rows: row0 + 32 * i

  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t0 = row0; t0 < t_hi; t0 += 32 * ADF_U) {
    float4 v[ADF_U];
#pragma unroll
    for (int u = 0; u < ADF_U; ++u) {
      const int t = t0 + 32 * u;
      v[u] = (cok && t < t_hi) ? adf_load4<XDT>(x, base + (long long)t * x_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < ADF_U; ++u) {
      const int t = t0 + 32 * u;
      if (t < t_hi) {
        *reinterpret_cast<float4*>(slab + (t - t_lo) * 32 + 4 * c4) = v[u];
        s4.x += v[u].x; s4.y += v[u].y; s4.z += v[u].z; s4.w += v[u].w;
      }
    }
  }
  // CTA-level reduction (4 row lanes of a warp, then 8 warps) into part[which]; then the cluster-level sum
  auto reduce_cluster = [&](float4 p, int which) -> float4 {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      p.x += __shfl_xor_sync(0xffffffffu, p.x, o); p.y += __shfl_xor_sync(0xffffffffu, p.y, o);
      p.z += __shfl_xor_sync(0xffffffffu, p.z, o); p.w += __shfl_xor_sync(0xffffffffu, p.w, o);
    }
    if (r4 == 0) *reinterpret_cast<float4*>(&red[w][4 * c4]) = p;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
      part[which][threadIdx.x] = t;
    }
    adf_cluster_sync();                                     // every CTA's partial is published (also a CTA barrier)
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (uint32_t r = 0; r < ADF_CS; ++r) {
      tot.x += adf_ld_peer(&part[which][4 * c4], r);     tot.y += adf_ld_peer(&part[which][4 * c4 + 1], r);
      tot.z += adf_ld_peer(&part[which][4 * c4 + 2], r); tot.w += adf_ld_peer(&part[which][4 * c4 + 3], r);
    }
    return tot;
  };
  const float inv_len = len > 0 ? 1.f / len : 0.f;
  float4 mean = reduce_cluster(s4, 0);
  mean.x *= inv_len; mean.y *= inv_len; mean.z *= inv_len; mean.w *= inv_len;
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = row0; t < t_hi; t += 32) {
    const float4 v = *reinterpret_cast<const float4*>(slab + (t - t_lo) * 32 + 4 * c4);
    const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
    q4.x += dx * dx; q4.y += dy * dy; q4.z += dz * dz; q4.w += dw * dw;
  }
  const float4 var = reduce_cluster(q4, 1);
  adf_cluster_sync();                                       // peers have read this CTA's partials: it may exit later
  if (!cok) return;
  const float4 rstd = make_float4(rsqrtf(var.x * inv_len + eps), rsqrtf(var.y * inv_len + eps),
                                  rsqrtf(var.z * inv_len + eps), rsqrtf(var.w * inv_len + eps));
  if (stats_out != nullptr && rank == 0 && threadIdx.x < 8) {
    float* so = stats_out + ((long long)b * C + c) * 2;
    so[0] = mean.x; so[1] = rstd.x; so[2] = mean.y; so[3] = rstd.y; so[4] = mean.z; so[5] = rstd.z; so[6] = mean.w; so[7] = rstd.w;
  }
  const float4 g = *reinterpret_cast<const float4*>(gb + (long long)b * gb_ld + c);
  const float4 be = *reinterpret_cast<const float4*>(gb + (long long)b * gb_ld + C + c);
  const float4 sc = make_float4(rstd.x * (1.f + g.x), rstd.y * (1.f + g.y), rstd.z * (1.f + g.z), rstd.w * (1.f + g.w));
  auto act4 = [&](float4 v) -> float4 {
    v.x = (v.x - mean.x) * sc.x + be.x; v.y = (v.y - mean.y) * sc.y + be.y;
    v.z = (v.z - mean.z) * sc.z + be.z; v.w = (v.w - mean.w) * sc.w + be.w;
    v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
    return v;
  };
  auto act_at = [&](int t) -> float4 {                      // t in this CTA's range, or its first row past it
    if (t >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < t_lo + TR) return act4(*reinterpret_cast<const float4*>(slab + (t - t_lo) * 32 + 4 * c4));
    return act4(adf_load4<XDT>(x, base + (long long)t * x_ld));   // the next CTA's first row (pool variant only)
  };
  const int t_end = min(T, t_lo + TR);                      // rows this CTA writes (zeros beyond len)
  if (!UP) {
#pragma unroll 4
    for (int t = row0; t < t_end; t += 32) st4any(out, ((long long)b * T + t) * out_ld + c, act_at(t), odt);
  } else {
    float w0[4], w1[4], w2[4], ub[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { w0[e] = up_w[(c + e) * 3]; w1[e] = up_w[(c + e) * 3 + 1]; w2[e] = up_w[(c + e) * 3 + 2]; ub[e] = up_b[c + e]; }
    for (int t = row0; t < t_end; t += 32) {
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), od = ev;
      if (t < len) {
        const float4 a0 = act_at(t), a1 = act_at(t + 1);
        ev = make_float4(a0.x * w1[0] + ub[0], a0.y * w1[1] + ub[1], a0.z * w1[2] + ub[2], a0.w * w1[3] + ub[3]);
        od = make_float4(a0.x * w2[0] + a1.x * w0[0] + ub[0], a0.y * w2[1] + a1.y * w0[1] + ub[1],
                         a0.z * w2[2] + a1.z * w0[2] + ub[2], a0.w * w2[3] + a1.w * w0[3] + ub[3]);
      }
      st4any(out, ((long long)b * 2 * T + 2 * t) * out_ld + c, ev, odt);
      st4any(out, ((long long)b * 2 * T + 2 * t + 1) * out_ld + c, od, odt);
    }
  }
}

// TMA variant for fp32 input and T <= 1760: the whole [T, 32-channel] slab of a CTA is fetched by a handful
// of bulk tensor copies issued by one thread (100 KB in flight per CTA at T = 800, no registers involved),
// so the read phase runs at HBM speed instead of at (loads in flight per thread) / latency.
constexpr int ADT_MAX_ROWS = 1760;

template <bool UP>
__global__ void __launch_bounds__(256)
adain_tma_kernel(const __grid_constant__ CUtensorMap tmx, int T, int BR, int nbox, int C,
                 const float* __restrict__ gb, long long gb_ld, float eps, float slope,
                 const int* __restrict__ lens, const float* __restrict__ up_w, const float* __restrict__ up_b,
                 void* out, int odt, long long out_ld, float* __restrict__ stats_out) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(128) float slab_raw[];
  __shared__ float red[8][32];
  __shared__ __align__(8) unsigned long long bar_storage;
  float* slab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(slab_raw) + 127) & ~uintptr_t(127));   // [nbox*BR][32]
  const uint32_t bar = smem_u32(&bar_storage);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r4 = lane >> 3, c4 = lane & 7;
  const int b = blockIdx.y, cb = blockIdx.x * 32, c = cb + 4 * c4;
  const bool cok = c < C;
  const int len = lens ? min(lens[b], T) : T;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)nbox * BR * 128u);
    for (int k = 0; k < nbox; ++k) tma_load_3d(smem_u32(slab) + (uint32_t)k * BR * 128u, &tmx, bar, cb, k * BR, b);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  const int row0 = w * 4 + r4;
  // explicit ld.shared: through the re-aligned generic pointer the compiler emits generic LD instead of LDS
  const uint32_t slab_u = smem_u32(slab) + 16u * c4;
  auto row4 = [&](int t) -> float4 {
    const uint4 r = lds128(slab_u + (uint32_t)t * 128u);
    return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
  };
  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = row0; t < len; t += 32) {
    const float4 v = row4(t);
    s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
  }
  auto reduce4 = [&](float4 p) -> float4 {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      p.x += __shfl_xor_sync(0xffffffffu, p.x, o); p.y += __shfl_xor_sync(0xffffffffu, p.y, o);
      p.z += __shfl_xor_sync(0xffffffffu, p.z, o); p.w += __shfl_xor_sync(0xffffffffu, p.w, o);
    }
    __syncthreads();
    if (r4 == 0) *reinterpret_cast<float4*>(&red[w][4 * c4]) = p;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(&red[i][4 * c4]);
      t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
    }
    return t;
  };
  const float inv_len = len > 0 ? 1.f / len : 0.f;
  float4 mean = reduce4(s4);
  mean.x *= inv_len; mean.y *= inv_len; mean.z *= inv_len; mean.w *= inv_len;
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = row0; t < len; t += 32) {
    const float4 v = row4(t);
    const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
    q4.x += dx * dx; q4.y += dy * dy; q4.z += dz * dz; q4.w += dw * dw;
  }
  const float4 var = reduce4(q4);
  if (!cok) return;
  const float4 rstd = make_float4(rsqrtf(var.x * inv_len + eps), rsqrtf(var.y * inv_len + eps),
                                  rsqrtf(var.z * inv_len + eps), rsqrtf(var.w * inv_len + eps));
  if (stats_out != nullptr && threadIdx.x < 8) {
    float* so = stats_out + ((long long)b * C + c) * 2;
    so[0] = mean.x; so[1] = rstd.x; so[2] = mean.y; so[3] = rstd.y; so[4] = mean.z; so[5] = rstd.z; so[6] = mean.w; so[7] = rstd.w;
  }
  const float4 g = *reinterpret_cast<const float4*>(gb + (long long)b * gb_ld + c);
  const float4 be = *reinterpret_cast<const float4*>(gb + (long long)b * gb_ld + C + c);
  const float4 sc = make_float4(rstd.x * (1.f + g.x), rstd.y * (1.f + g.y), rstd.z * (1.f + g.z), rstd.w * (1.f + g.w));
  auto act_at = [&](int t) -> float4 {
    if (t >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = row4(t);
    v.x = (v.x - mean.x) * sc.x + be.x; v.y = (v.y - mean.y) * sc.y + be.y;
    v.z = (v.z - mean.z) * sc.z + be.z; v.w = (v.w - mean.w) * sc.w + be.w;
    v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
    return v;
  };
  if (!UP) {
#pragma unroll 4
    for (int t = row0; t < T; t += 32) st4any(out, ((long long)b * T + t) * out_ld + c, act_at(t), odt);
  } else {
    float w0[4], w1[4], w2[4], ub[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { w0[e] = up_w[(c + e) * 3]; w1[e] = up_w[(c + e) * 3 + 1]; w2[e] = up_w[(c + e) * 3 + 2]; ub[e] = up_b[c + e]; }
    for (int t = row0; t < T; t += 32) {
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), od = ev;
      if (t < len) {
        const float4 a0 = act_at(t), a1 = act_at(t + 1);
        ev = make_float4(a0.x * w1[0] + ub[0], a0.y * w1[1] + ub[1], a0.z * w1[2] + ub[2], a0.w * w1[3] + ub[3]);
        od = make_float4(a0.x * w2[0] + a1.x * w0[0] + ub[0], a0.y * w2[1] + a1.y * w0[1] + ub[1],
                         a0.z * w2[2] + a1.z * w0[2] + ub[2], a0.w * w2[3] + a1.w * w0[3] + ub[3]);
      }
      st4any(out, ((long long)b * 2 * T + 2 * t) * out_ld + c, ev, odt);
      st4any(out, ((long long)b * 2 * T + 2 * t + 1) * out_ld + c, od, odt);
    }
  }
}

// Persistent ring variant (the one the decoder's shapes take).  ncu on the kernel above: DRAM is busy 19 % of the
// time -- the pass is bound by INSTRUCTION ISSUE, not memory (about 29 instructions per element over three
// shared-memory passes, block-wide reductions in every thread, loads / statistics / stores of the resident CTAs in
// lockstep); neither a contiguous-slab layout, nor 512-byte rows split over a cluster, nor cp.async instead of TMA
// changed the 2.6-2.9 TB/s.  This version (a) makes ONE persistent CTA per SM walk the (item, 32-channel slab)
// units through an S-stage shared-memory ring fed by TMA, so the copies of units k+1 .. k+S-1 cost no instructions
// and are in flight while unit k is processed, and (b) cuts the arithmetic to two passes: one statistics pass
// (sums of d = x - pivot and d^2, pivot = the channel's first frame, so no digits are lost on bias-dominated
// channels), a 64-thread fold of the warps' partials, and one fused-multiply-add + max + pack + 8-byte store per
// four elements.
constexpr int ADR_THREADS = 512;    // 1024 threads (64 registers) measured slower: 23.5 vs 20.0 us at 16x800x1024
constexpr int ADR_WARPS = ADR_THREADS / 32;
constexpr int ADR_MAX_STAGES = 4;

// MX: 0 <= slope <= 1, LeakyReLU(y) = max(y, slope * y) -- a template parameter so that the unrolled loops are branch-free
template <bool UP, bool MX, int W>
__global__ void __launch_bounds__(ADR_THREADS, 1)
adain_ring_kernel(const __grid_constant__ CUtensorMap tmx, int T, int BR, int nbox, int C, int nslab, int units, int S,
                  const float* __restrict__ gb, long long gb_ld, float eps, float slope,
                  const int* __restrict__ lens, const float* __restrict__ up_w, const float* __restrict__ up_b,
                  void* out, int odt, long long out_ld, float* __restrict__ stats_out) {
  extern __shared__ __align__(128) float slab_raw[];
  // W = channels per slab (32, or 16 when that fills the SMs' last round better: see as_adain_norm_apply)
  constexpr int CG = W / 4;                                 // 4-channel column groups per row
  constexpr int RPW = 32 / CG;                              // rows a warp covers per pass
  constexpr int ROWS = ADR_WARPS * RPW;                     // rows covered per pass iteration
  constexpr uint32_t ROWB = 4u * W;                         // bytes per slab row
  __shared__ __align__(16) float red[2][ADR_WARPS][W];      // per-warp partial sums / squared sums
  __shared__ __align__(16) float tot[2][W];
  __shared__ __align__(8) unsigned long long bars[ADR_MAX_STAGES];
  float* ring = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(slab_raw) + 127) & ~uintptr_t(127));   // [S][nbox*BR][32]
  const uint32_t slab_bytes = (uint32_t)nbox * BR * ROWB;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r4 = lane / CG, c4 = lane % CG;
  const int row0 = w * RPW + r4;
  auto issue = [&](int u, int stage) {                      // one thread: nbox bulk tensor copies, rows >= T / channels >= C arrive as zeros
    const uint32_t bar = smem_u32(&bars[stage]);
    const uint32_t dst = smem_u32(ring) + (uint32_t)stage * slab_bytes;
    const int b = u / nslab, cb = (u - b * nslab) * W;
    mbar_expect_tx(bar, slab_bytes);
    for (int k = 0; k < nbox; ++k) tma_load_3d(dst + (uint32_t)k * BR * ROWB, &tmx, bar, cb, k * BR, b);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      const int u = blockIdx.x + s * gridDim.x;
      if (u < units) issue(u, s);
    }
  }
  __syncthreads();
  int stage = 0;
  uint32_t phase = 0;
  // per-unit scalars (length, gamma, beta) are fetched one unit ahead with volatile loads: issued at the top of unit
  // k for unit k + 1, consumed a whole unit later, so their L2 / DRAM latency never stalls the in-order pipeline
  // (ncu: with the loads next to their use, 30 % of the samples sat on the first instruction that reads gamma)
  auto load_scalars = [&](int u, int& len_o, float4& g_o, float4& be_o) {
    const int b = u / nslab, c = (u - b * nslab) * W + 4 * c4;
    len_o = T;
    if (lens) { asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(len_o) : "l"(lens + b)); }
    if (c < C) {
      const float* gp = gb + (long long)b * gb_ld + c;
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g_o.x), "=f"(g_o.y), "=f"(g_o.z), "=f"(g_o.w) : "l"(gp));
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(be_o.x), "=f"(be_o.y), "=f"(be_o.z), "=f"(be_o.w) : "l"(gp + C));
    }
  };
  int len_n = T;
  float4 g_n = make_float4(0.f, 0.f, 0.f, 0.f), be_n = g_n;
  if ((int)blockIdx.x < units) load_scalars(blockIdx.x, len_n, g_n, be_n);
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = u / nslab, cb = (u - b * nslab) * W, c = cb + 4 * c4;
    const bool cok = c < C;
    int len = len_n;
    const float4 g = g_n, be = be_n;
    if (u + (int)gridDim.x < units) load_scalars(u + gridDim.x, len_n, g_n, be_n);
    mbar_wait(smem_u32(&bars[stage]), phase);
    len = min(len, T);
    // explicit ld.shared: through the re-aligned generic pointer the compiler emits generic LD instead of LDS
    const uint32_t slab = smem_u32(ring) + (uint32_t)stage * slab_bytes + 16u * c4;
    auto row4 = [&](int t) -> float4 {
      const uint4 r = lds128(slab + (uint32_t)t * ROWB);
      return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
    };
    const float4 pv = row4(0);
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = s4;
#pragma unroll 4
    for (int t = row0; t < len; t += ROWS) {
      const float4 v = row4(t);
      const float dx = v.x - pv.x, dy = v.y - pv.y, dz = v.z - pv.z, dw = v.w - pv.w;
      s4.x += dx; s4.y += dy; s4.z += dz; s4.w += dw;
      q4.x = fmaf(dx, dx, q4.x); q4.y = fmaf(dy, dy, q4.y); q4.z = fmaf(dz, dz, q4.z); q4.w = fmaf(dw, dw, q4.w);
    }
#pragma unroll
    for (int o = CG; o <= 16; o <<= 1) {
      s4.x += __shfl_xor_sync(0xffffffffu, s4.x, o); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, o);
      s4.z += __shfl_xor_sync(0xffffffffu, s4.z, o); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, o);
      q4.x += __shfl_xor_sync(0xffffffffu, q4.x, o); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, o);
      q4.z += __shfl_xor_sync(0xffffffffu, q4.z, o); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, o);
    }
    if (r4 == 0) {
      *reinterpret_cast<float4*>(&red[0][w][4 * c4]) = s4;
      *reinterpret_cast<float4*>(&red[1][w][4 * c4]) = q4;
    }
    __syncthreads();
    if (threadIdx.x < 2 * W) {                              // 2 x W columns: one thread folds the 16 warps' partials
      const int which = threadIdx.x / W, col = threadIdx.x % W;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < ADR_WARPS; ++i) a += red[which][i][col];
      tot[which][col] = a;
    }
    __syncthreads();
    if (cok) {
      const float inv_len = len > 0 ? 1.f / len : 0.f;
      const float4 sd = *reinterpret_cast<const float4*>(&tot[0][4 * c4]);
      const float4 sq = *reinterpret_cast<const float4*>(&tot[1][4 * c4]);
      const float4 md = make_float4(sd.x * inv_len, sd.y * inv_len, sd.z * inv_len, sd.w * inv_len);   // mean - pivot
      const float4 mean = make_float4(pv.x + md.x, pv.y + md.y, pv.z + md.z, pv.w + md.w);
      const float4 rstd = make_float4(rsqrtf(fmaxf(sq.x * inv_len - md.x * md.x, 0.f) + eps), rsqrtf(fmaxf(sq.y * inv_len - md.y * md.y, 0.f) + eps),
                                      rsqrtf(fmaxf(sq.z * inv_len - md.z * md.z, 0.f) + eps), rsqrtf(fmaxf(sq.w * inv_len - md.w * md.w, 0.f) + eps));
      if (stats_out != nullptr && threadIdx.x < CG) {
        float* so = stats_out + ((long long)b * C + c) * 2;
        so[0] = mean.x; so[1] = rstd.x; so[2] = mean.y; so[3] = rstd.y; so[4] = mean.z; so[5] = rstd.z; so[6] = mean.w; so[7] = rstd.w;
      }
      // y = (x - mean) * rstd * (1 + g) + be = x * sc + sh: one fused multiply-add per element (its single rounding
      // is relative to |x * sc| <= ~|mean| / std, i.e. ~1e-5 of the normalised scale for the worst decoder channels)
      const float4 sc = make_float4(rstd.x * (1.f + g.x), rstd.y * (1.f + g.y), rstd.z * (1.f + g.z), rstd.w * (1.f + g.w));
      const float4 sh = make_float4(be.x - mean.x * sc.x, be.y - mean.y * sc.y, be.z - mean.z * sc.z, be.w - mean.w * sc.w);
      auto act_row = [&](int t) -> float4 {                 // t < len
        float4 v = row4(t);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (MX) {
          v.x = fmaxf(v.x, v.x * slope); v.y = fmaxf(v.y, v.y * slope); v.z = fmaxf(v.z, v.z * slope); v.w = fmaxf(v.w, v.w * slope);
        } else {
          v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
          v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        }
        return v;
      };
      if (!UP) {
        if (odt != AS_F32) {
          // 16-bit output: 8 bytes per thread, 64 contiguous bytes per row; frames past the length are zeros
          uint2* o2 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + ((long long)b * T + row0) * out_ld + c);
          const long long ostep = (long long)ROWS * out_ld / 4;     // in uint2 units (out_ld % 4 == 0)
          int t = row0;
          if (odt == AS_F16) {
#pragma unroll 4
            for (; t < len; t += ROWS, o2 += ostep) { const float4 v = act_row(t); *o2 = make_uint2(pack16(v.x, v.y, AS_F16), pack16(v.z, v.w, AS_F16)); }
          } else {
#pragma unroll 4
            for (; t < len; t += ROWS, o2 += ostep) { const float4 v = act_row(t); *o2 = make_uint2(pack16(v.x, v.y, AS_BF16), pack16(v.z, v.w, AS_BF16)); }
          }
          for (; t < T; t += ROWS, o2 += ostep) *o2 = make_uint2(0u, 0u);
        } else {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
          for (int t = row0; t < T; t += ROWS) st4any(out, ((long long)b * T + t) * out_ld + c, t < len ? act_row(t) : z4, odt);
        }
      } else {
        float w0[4], w1[4], w2[4], ub[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { w0[e] = up_w[(c + e) * 3]; w1[e] = up_w[(c + e) * 3 + 1]; w2[e] = up_w[(c + e) * 3 + 2]; ub[e] = up_b[c + e]; }
        for (int t = row0; t < T; t += ROWS) {
          float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), od = ev;
          if (t < len) {
            const float4 a0 = act_row(t), a1 = t + 1 < len ? act_row(t + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            ev = make_float4(a0.x * w1[0] + ub[0], a0.y * w1[1] + ub[1], a0.z * w1[2] + ub[2], a0.w * w1[3] + ub[3]);
            od = make_float4(a0.x * w2[0] + a1.x * w0[0] + ub[0], a0.y * w2[1] + a1.y * w0[1] + ub[1],
                             a0.z * w2[2] + a1.z * w0[2] + ub[2], a0.w * w2[3] + a1.w * w0[3] + ub[3]);
          }
          st4any(out, ((long long)b * 2 * T + 2 * t) * out_ld + c, ev, odt);
          st4any(out, ((long long)b * 2 * T + 2 * t + 1) * out_ld + c, od, odt);
        }
      }
    }
    __syncthreads();   // every thread is done with this stage (and with red / tot): refill it
    if (threadIdx.x == 0) {
      const int un = u + S * gridDim.x;
      if (un < units) issue(un, stage);
    }
    if (++stage == S) { stage = 0; phase ^= 1u; }
  }
}

// ---------------------------------------------------------------------------------------------
// expansion of token columns along a monotonic alignment: out[b, c, y] = x[b, c, tok[b, y]]
// (= x @ path for a 0/1 path with one 1 per column).  One warp-row per (b, c): writes coalesced along y,
// reads walk the row of x monotonically (every 128-byte line is fetched once and re-served by L1).
// ---------------------------------------------------------------------------------------------
__global__ void expand_tokens_kernel(const float* __restrict__ x, const int* __restrict__ tok, float* __restrict__ out,
                                     long long rows, int C, int Tx, int Ty) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = rows * Ty;
  ASB_GRID_STRIDE(i, total, {
    const int y = i % Ty;
    const long long bc = i / Ty;
    const int t = tok[(bc / C) * Ty + y];
    out[i] = (t >= 0 && t < Tx) ? x[bc * Tx + t] : 0.f;
  })
}

// ---------------------------------------------------------------------------------------------
// nearest upsample along T
// ---------------------------------------------------------------------------------------------
__global__ void repeat_rows_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                   int C, int rep, const int* __restrict__ lens, void* out, int odt,
                                   long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const int To = T * rep;
  const long long total = (long long)B * To * C;
  ASB_GRID_STRIDE(i, total, {
    const int c = i % C;
    const auto r = i / C;
    const int to = r % To;
    const int b = r / To;
    const int t = to / rep;
    const int len = lens ? min(lens[b], T) : T;
    float v = t < len ? ldany(x, ((long long)b * T + t) * x_ld + c, xdt) : 0.f;
    stany(out, ((long long)b * To + to) * out_ld + c, v, odt);
  })
}

// ---------------------------------------------------------------------------------------------
// length regulation: one block per item builds frame->token map by a prefix sum, then gathers
// ---------------------------------------------------------------------------------------------
constexpr int LR_ROWS = 32;   // output rows per CTA

__global__ void length_regulate_kernel(const void* __restrict__ x, int xdt, long long x_ld, int Tt,
                                       int C, const int* __restrict__ dur,
                                       const int* __restrict__ lens_t, int rep, int To, void* out,
                                       int odt, long long out_ld, int* __restrict__ out_lens, int vec) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ int cum[];  // [Tt + 1] inclusive prefix sums, cum[0] = 0
  const int b = blockIdx.y;
  const int nt = lens_t ? min(lens_t[b], Tt) : Tt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (warp == 0) {
    // chunked warp scan of the durations (every CTA of the item recomputes it: Tt is a few hundred)
    int carry = 0;
    if (lane == 0) cum[0] = 0;
    for (int j0 = 0; j0 < nt; j0 += 32) {
      const int j = j0 + lane;
      int v = j < nt ? max(dur[(long long)b * Tt + j], 0) : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      if (j < nt) cum[j + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  const int L = cum[nt];
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_lens) out_lens[b] = min(L * rep, To);
  const int r_end = min(To, (int)(blockIdx.x + 1) * LR_ROWS);
  for (int r = blockIdx.x * LR_ROWS + warp; r < r_end; r += nw) {
    const int fr = r / rep;
    int tok = -1;
    if (fr < L) {
      int lo = 0, hi = nt;  // find largest j with cum[j] <= fr
      while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cum[mid] <= fr) lo = mid; else hi = mid; }
      tok = lo;
    }
    const long long xo = ((long long)b * Tt + max(tok, 0)) * x_ld, yo = ((long long)b * To + r) * out_ld;
    if (vec == 1) {          // same element size on both sides, 16-byte aligned rows: straight 16-byte copies
      const int n16 = C * (xdt == AS_F32 ? 4 : 2) / 16;
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(x) + xo * (xdt == AS_F32 ? 4 : 2));
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + yo * (odt == AS_F32 ? 4 : 2));
      for (int c = lane; c < n16; c += 32) dst[c] = tok >= 0 ? __ldg(src + c) : make_uint4(0u, 0u, 0u, 0u);
    } else if (vec == 2) {   // fp32 -> 16-bit, 8 channels per lane
      const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + xo);
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + yo);
      for (int c = lane; c < C / 8; c += 32) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (tok >= 0) {
          const float4 p = __ldg(src + 2 * c), q = __ldg(src + 2 * c + 1);
          o = make_uint4(pack16(p.x, p.y, odt), pack16(p.z, p.w, odt), pack16(q.x, q.y, odt), pack16(q.z, q.w, odt));
        }
        dst[c] = o;
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        float v = tok >= 0 ? ldany(x, xo + c, xdt) : 0.f;
        stany(out, yo + c, v, odt);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// direct convolution, tiny Cin
// ---------------------------------------------------------------------------------------------
constexpr int CS_MAX_TAPS = 32;
struct SmallTaps { int dt[CS_MAX_TAPS]; int df[CS_MAX_TAPS]; };

__global__ void conv_small_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                  int F, int Cin, const float* __restrict__ w,
                                  const float* __restrict__ bias, int ntaps, SmallTaps taps,
                                  int Cout, const int* __restrict__ lens, void* yr, int yrdt,
                                  long long yr_ld, void* ya, int yadt, long long ya_ld, int act,
                                  float slope) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = (long long)B * T * F * Cout;
  ASB_GRID_STRIDE(i, total, {
    const int co = i % Cout;
    const auto row = i / Cout;
    const int f = row % F;
    const int t = (row / F) % T;
    const int b = (row / F) / T;
    float acc = bias ? bias[co] : 0.f;
    for (int j = 0; j < ntaps; ++j) {
      const int ti = t + taps.dt[j], fi = f + taps.df[j];
      if (ti < 0 || ti >= T || fi < 0 || fi >= F) continue;
      const long long xr = (((long long)b * T + ti) * F + fi) * x_ld;
      const float* wj = w + ((long long)j * Cout + co) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc += ldany(x, xr + ci, xdt) * wj[ci];
    }
    if (lens && t >= lens[b]) acc = 0.f;
    if (yr) stany(yr, row * yr_ld + co, acc, yrdt);
    if (ya) stany(ya, row * ya_ld + co, apply_act(acc, act, slope), yadt);
  })
}

// 8 output channels per thread, weights transposed to [tap][ci][co] in shared memory, 32-bit index
// arithmetic (the scalar kernel above spends its time in 64-bit div/mod: 530 us for the 1 -> 64
// channel 3x3 stems of JDCNet / Mel_block on a 16 x 240 x 80 image).
// P consecutive F positions per thread (P = 4 when F % 4 == 0): a weight octet read from shared memory and the
// row / bounds arithmetic of a tap are shared by the P positions.
template <int P>
__global__ void __launch_bounds__(256)
conv_small_vec_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F, int Cin,
                      const float* __restrict__ w, const float* __restrict__ bias, int ntaps, SmallTaps taps,
                      int Cout, const int* __restrict__ lens, void* yr, int yrdt, long long yr_ld, void* ya,
                      int yadt, long long ya_ld, int act, float slope, int vec_raw, int vec_act) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ float ws[];   // [ntaps][Cin][Cout] + bias[Cout]
  const int nw = ntaps * Cin * Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    const int co = i % Cout, ci = (i / Cout) % Cin, j = i / (Cout * Cin);
    ws[i] = w[((long long)j * Cout + co) * Cin + ci];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[nw + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout >> 3;
  const unsigned rows = (unsigned)B * T * F;            // host guarantees rows * groups < 2^31
  const unsigned total = rows / P * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups, row = i / groups * P;
    const int f = row % F;
    const unsigned bt = row / F;
    const int t = bt % T, b = bt / T;
    const int co = g * 8;
    float acc[P][8];
#pragma unroll
    for (int q = 0; q < P; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[q][e] = ws[nw + co + e];
    for (int j = 0; j < ntaps; ++j) {
      const int ti = t + taps.dt[j], f0 = f + taps.df[j];
      if (ti < 0 || ti >= T) continue;
      const long long xr = (((long long)b * T + ti) * F + f0) * x_ld;
      for (int ci = 0; ci < Cin; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (j * Cin + ci) * Cout + co);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (j * Cin + ci) * Cout + co + 4);
#pragma unroll
        for (int q = 0; q < P; ++q) {
          const int fi = f0 + q;
          const float xv = (fi >= 0 && fi < F) ? ldany(x, xr + q * x_ld + ci, xdt) : 0.f;
          acc[q][0] += xv * w0.x; acc[q][1] += xv * w0.y; acc[q][2] += xv * w0.z; acc[q][3] += xv * w0.w;
          acc[q][4] += xv * w1.x; acc[q][5] += xv * w1.y; acc[q][6] += xv * w1.z; acc[q][7] += xv * w1.w;
        }
      }
    }
    const bool dead = lens && t >= lens[b];
    auto put = [&](void* y, int dt, long long ld, int vec, unsigned r, const float (&v)[8]) {
      const long long o = (long long)r * ld + co;
      if (vec && dt != AS_F32) {
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y) + o) =
            make_uint4(pack16(v[0], v[1], dt), pack16(v[2], v[3], dt), pack16(v[4], v[5], dt), pack16(v[6], v[7], dt));
      } else if (vec) {
        float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + o);
        q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) stany(y, o + e, v[e], dt);
      }
    };
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (dead) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
      }
      if (yr) put(yr, yrdt, yr_ld, vec_raw, row + q, acc[q]);
      if (ya) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = apply_act(acc[q][e], act, slope);
        put(ya, yadt, ya_ld, vec_act, row + q, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise convolution with stride (optionally GLU on a 2C-channel input)
// ---------------------------------------------------------------------------------------------
__global__ void dwconv_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F,
                              int C, int glu, const float* __restrict__ w,
                              const float* __restrict__ bias, int kt, int kf, int st, int sf, int pt,
                              int pf, int To, int Fo, const int* __restrict__ lens_in,
                              const int* __restrict__ lens_out, int act, float slope, void* out,
                              int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = (long long)B * To * Fo * C;
  ASB_GRID_STRIDE(i, total, {
    const int c = i % C;
    const auto row = i / C;
    const int fo = row % Fo;
    const int to = (row / Fo) % To;
    const int b = (row / Fo) / To;
    const int len_in = lens_in ? min(lens_in[b], T) : T;
    float acc = bias ? bias[c] : 0.f;
    for (int jt = 0; jt < kt; ++jt) {
      const int ti = to * st + jt - pt;
      if (ti < 0 || ti >= len_in) continue;
      for (int jf = 0; jf < kf; ++jf) {
        const int fi = fo * sf + jf - pf;
        if (fi < 0 || fi >= F) continue;
        const long long xr = (((long long)b * T + ti) * F + fi) * x_ld;
        float v = ldany(x, xr + c, xdt);
        if (glu) { float g = ldany(x, xr + C + c, xdt); v = v / (1.f + __expf(-g)); }
        acc += v * w[(jt * kf + jf) * C + c];
      }
    }
    float y = apply_act(acc, act, slope);
    if (lens_out && to >= lens_out[b]) y = 0.f;
    stany(out, row * out_ld + c, y, odt);
  })
}

// Time-only depthwise convolution with stride 1 (the conformer's k = 31 module, convolution.py:136-149):
// the generic kernel recomputes the GLU of every input 31 times and reads it from global memory per
// tap (104 us under ncu).  Here a CTA stages a (64 + k - 1) x 64-channel tile of GLU'd inputs in
// shared memory once; each thread keeps its channel's taps in registers.
constexpr int DWT_ROWS = 64, DWT_CH = 64, DWT_MAXK = 32;

__global__ void __launch_bounds__(256)
dwconv_time_tiled_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T, int C, int glu,
                         const float* __restrict__ w, const float* __restrict__ bias, int kt, int pt,
                         const int* __restrict__ lens_in, const int* __restrict__ lens_out, int act,
                         float slope, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ float tile[];   // [DWT_ROWS + kt - 1][DWT_CH]
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int b = blockIdx.z, t0 = blockIdx.y * DWT_ROWS, c = blockIdx.x * DWT_CH + tx;
  const int len_in = lens_in ? min(lens_in[b], T) : T;
  const int nrows = DWT_ROWS + kt - 1;
  const bool cok = c < C;
  // four rows per iteration, all loads issued before the first use (one row at a time the ~24 dependent pairs of
  // loads per thread set the kernel's time)
  for (int r0 = ty; r0 < nrows; r0 += 16) {
    float v[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ti = t0 - pt + r0 + 4 * k;
      const bool ok = cok && r0 + 4 * k < nrows && ti >= 0 && ti < len_in;
      const long long xr = ((long long)b * T + (ok ? ti : 0)) * x_ld;
      v[k] = ok ? ldany(x, xr + c, xdt) : 0.f;
      g[k] = (ok && glu) ? ldany(x, xr + C + c, xdt) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + 4 * k;
      if (r < nrows) tile[r * DWT_CH + tx] = glu ? v[k] / (1.f + __expf(-g[k])) : v[k];
    }
  }
  float wr[DWT_MAXK];
#pragma unroll
  for (int j = 0; j < DWT_MAXK; ++j) wr[j] = (cok && j < kt) ? w[j * C + c] : 0.f;
  const float bv = (cok && bias) ? bias[c] : 0.f;
  __syncthreads();
  if (!cok) return;
  const int len_out = lens_out ? lens_out[b] : T;
  for (int r = ty; r < DWT_ROWS; r += 4) {
    const int to = t0 + r;
    if (to >= T) break;
    float acc = bv;
#pragma unroll
    for (int j = 0; j < DWT_MAXK; ++j)
      if (j < kt) acc += tile[(r + j) * DWT_CH + tx] * wr[j];
    float y = apply_act(acc, act, slope);
    if (to >= len_out) y = 0.f;
    stany(out, ((long long)b * T + to) * out_ld + c, y, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// average pooling with replicate padding of the last T position
// ---------------------------------------------------------------------------------------------
__global__ void avgpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F,
                               int C, int pt, int pf, int To, int Fo, void* out, int odt,
                               long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = (long long)B * To * Fo * C;
  const float inv = 1.f / (pt * pf);
  ASB_GRID_STRIDE(i, total, {
    const int c = i % C;
    const auto row = i / C;
    const int fo = row % Fo;
    const int to = (row / Fo) % To;
    const int b = (row / Fo) / To;
    float acc = 0.f;
    for (int jt = 0; jt < pt; ++jt) {
      const int ti = min(to * pt + jt, T - 1);  // replicate the last column when T is odd
      for (int jf = 0; jf < pf; ++jf) {
        const int fi = fo * pf + jf;
        acc += ldany(x, (((long long)b * T + ti) * F + fi) * x_ld + c, xdt);
      }
    }
    stany(out, row * out_ld + c, acc * inv, odt);
  })
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(eval) affine -> LeakyReLU -> MaxPool along F
// ---------------------------------------------------------------------------------------------
__global__ void affine_act_maxpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B,
                                          int T, int F, int C, const float* __restrict__ scale,
                                          const float* __restrict__ shift, float slope, int pf, int Fo,
                                          void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = (long long)B * T * Fo * C;
  ASB_GRID_STRIDE(i, total, {
    const int c = i % C;
    const auto row = i / C;
    const int fo = row % Fo;
    const auto bt = row / Fo;
    const float sc = scale[c], sh = shift[c];
    float m = -INFINITY;
    for (int j = 0; j < pf; ++j) {
      float v = ldany(x, (bt * F + fo * pf + j) * x_ld + c, xdt) * sc + sh;
      v = v > 0.f ? v : v * slope;
      m = fmaxf(m, v);
    }
    stany(out, row * out_ld + c, m, odt);
  })
}

// ---------------------------------------------------------------------------------------------
// 8-channel (16-byte) variants of the strided depthwise conv / average pool / affine+LeakyReLU+max pool for 16-bit
// channels-last tensors: the scalar kernels above issue one 2-byte load and a chain of 64-bit div/mods per
// element (dwconv 120 us, pools 47 us on the 16 x 240 x 80 x 64 style-encoder images -- 10x off the HBM time).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8h(const void* p, long long off, int dt, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + off));
  const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (dt == AS_BF16) {
      v[2 * e] = __uint_as_float(q[e] << 16); v[2 * e + 1] = __uint_as_float(q[e] & 0xFFFF0000u);
    } else {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[e]));
      v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void st8h(void* p, long long off, const float (&v)[8], int dt) {
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p) + off) =
      make_uint4(pack16(v[0], v[1], dt), pack16(v[2], v[3], dt), pack16(v[4], v[5], dt), pack16(v[6], v[7], dt));
}

__global__ void __launch_bounds__(256)
dwconv_vec8_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F, int C,
                   const float* __restrict__ w, const float* __restrict__ bias, int kt, int kf, int st, int sf, int pt,
                   int pf, int To, int Fo, const int* __restrict__ lens_in, const int* __restrict__ lens_out, int act,
                   float slope, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const unsigned groups = (unsigned)C >> 3;
  const unsigned total = (unsigned)B * To * Fo * groups;     // host guarantees < 2^31
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups, row = i / groups;
    const int fo = row % Fo;
    const unsigned bto = row / Fo;
    const int to = bto % To, b = bto / To;
    const int c = g * 8;
    const int len_in = lens_in ? min(lens_in[b], T) : T;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[c + e] : 0.f;
    for (int jt = 0; jt < kt; ++jt) {
      const int ti = to * st + jt - pt;
      if (ti < 0 || ti >= len_in) continue;
      for (int jf = 0; jf < kf; ++jf) {
        const int fi = fo * sf + jf - pf;
        if (fi < 0 || fi >= F) continue;
        float v[8];
        ld8h(x, (((long long)b * T + ti) * F + fi) * x_ld + c, xdt, v);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (jt * kf + jf) * C + c));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (jt * kf + jf) * C + c + 4));
        acc[0] += v[0] * w0.x; acc[1] += v[1] * w0.y; acc[2] += v[2] * w0.z; acc[3] += v[3] * w0.w;
        acc[4] += v[4] * w1.x; acc[5] += v[5] * w1.y; acc[6] += v[6] * w1.z; acc[7] += v[7] * w1.w;
      }
    }
    const bool dead = lens_out && to >= lens_out[b];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = dead ? 0.f : apply_act(acc[e], act, slope);
    st8h(out, (long long)row * out_ld + c, acc, odt);
  }
}

__global__ void __launch_bounds__(256)
avgpool_vec8_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F, int C, int pt, int pf,
                    int To, int Fo, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const unsigned groups = (unsigned)C >> 3;
  const unsigned total = (unsigned)B * To * Fo * groups;
  const float inv = 1.f / (pt * pf);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups, row = i / groups;
    const int fo = row % Fo;
    const unsigned bto = row / Fo;
    const int to = bto % To, b = bto / To;
    const int c = g * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int jt = 0; jt < pt; ++jt) {
      const int ti = min(to * pt + jt, T - 1);  // replicate the last column when T is odd
      for (int jf = 0; jf < pf; ++jf) {
        float v[8];
        ld8h(x, (((long long)b * T + ti) * F + fo * pf + jf) * x_ld + c, xdt, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    st8h(out, (long long)row * out_ld + c, acc, odt);
  }
}

__global__ void __launch_bounds__(256)
affine_act_maxpool_vec8_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F, int C,
                               const float* __restrict__ scale, const float* __restrict__ shift, float slope, int pf,
                               int Fo, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const unsigned groups = (unsigned)C >> 3;
  const unsigned total = (unsigned)B * T * Fo * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups, row = i / groups;
    const int fo = row % Fo;
    const unsigned bt = row / Fo;
    const int c = g * 8;
    float sc[8], sh[8], m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = scale[c + e]; sh[e] = shift[c + e]; m[e] = -INFINITY; }
    for (int j = 0; j < pf; ++j) {
      float v[8];
      ld8h(x, ((long long)bt * F + fo * pf + j) * x_ld + c, xdt, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float y = v[e] * sc[e] + sh[e];
        y = y > 0.f ? y : y * slope;
        m[e] = fmaxf(m[e], y);
      }
    }
    st8h(out, (long long)row * out_ld + c, m, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// LeakyReLU + global average pool over (T with stride, F)
// ---------------------------------------------------------------------------------------------
__global__ void global_avgpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T, int F,
                                      int C, int ts, float slope, void* out, int odt,
                                      long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  int n = 0;
  for (int t = 0; t < T; t += ts)
    for (int f = 0; f < F; ++f) {
      float v = ldany(x, (((long long)b * T + t) * F + f) * x_ld + c, xdt);
      acc += v > 0.f ? v : v * slope;
      ++n;
    }
  stany(out, (long long)b * out_ld + c, acc / n, odt);
}

// ---------------------------------------------------------------------------------------------
// one LSTM step from zero state (EMA_Predictor quirk)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void lstm_onestep_kernel(const float* __restrict__ xp, long long xp_ld, long long rows, int H,
                                    void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long total = rows * 2 * H;
  ASB_GRID_STRIDE(i, total, {
    const int j = i % H;
    const int dir = (i / H) % 2;
    const auto row = i / (2 * H);
    const float* g = xp + row * xp_ld + (long long)dir * 4 * H;
    const float ig = sigmoidf_(g[j]), gg = tanhf(g[2 * H + j]), og = sigmoidf_(g[3 * H + j]);
    const float c = ig * gg;
    stany(out, row * out_ld + dir * H + j, og * tanhf(c), odt);
  })
}

// ---------------------------------------------------------------------------------------------
// log-norm energy: mel fp32 [B, n_mels, T] -> [B, T]
// ---------------------------------------------------------------------------------------------
__global__ void log_norm_kernel(const float* __restrict__ mel, int B, int M, int T, float* out) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int t = i % T, b = i / T;
  float s = 0.f;
  for (int m = 0; m < M; ++m) {
    float e = expf(mel[((long long)b * M + m) * T + t] * 4.f - 4.f);
    s += e * e;
  }
  out[i] = logf(sqrtf(s));
}

// ---------------------------------------------------------------------------------------------
// channels-first <-> channels-last with cast / affine / masking (tiled transpose)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_cast_kernel(const void* __restrict__ src, int sdt, void* dst, int ddt, int C,
                                      int T, long long cl_ld, int to_cl, const float* __restrict__ sub,
                                      const float* __restrict__ mul, const int* __restrict__ lens) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int len = lens ? min(lens[b], T) : T;
  if (to_cl) {
    // read [c][t] coalesced along t, write [t][c] coalesced along c
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, t = t0 + threadIdx.x;
      float v = 0.f;
      if (c < C && t < T) {
        v = ldany(src, ((long long)b * C + c) * T + t, sdt);
        if (sub) v = (v - sub[c]) * mul[c];
        if (t >= len) v = 0.f;
      }
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int t = t0 + i, c = c0 + threadIdx.x;
      if (c < C && t < T) stany(dst, ((long long)b * T + t) * cl_ld + c, tile[threadIdx.x][i], ddt);
    }
  } else {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int t = t0 + i, c = c0 + threadIdx.x;
      float v = 0.f;
      if (c < C && t < T) {
        v = ldany(src, ((long long)b * T + t) * cl_ld + c, sdt);
        if (sub) v = (v - sub[c]) * mul[c];
        if (t >= len) v = 0.f;
      }
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, t = t0 + threadIdx.x;
      if (c < C && t < T) stany(dst, ((long long)b * C + c) * T + t, tile[threadIdx.x][i], ddt);
    }
  }
}

// 16-bit in / out, 8-channel groups on 16-byte boundaries, 32-bit element counts
static inline bool vec8_ok(const void* x, int xdt, long long x_ld, const void* out, int odt, long long out_ld, int C, long long total) {
  return xdt != AS_F32 && odt != AS_F32 && (C % 8) == 0 && (x_ld % 8) == 0 && (out_ld % 8) == 0 &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && total < (1ll << 31);
}

static inline unsigned ew_grid(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;  // a few waves of the 148 SMs, grid-stride beyond that
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

static inline bool dt_ok(int d) { return d == AS_F16 || d == AS_BF16 || d == AS_F32; }

}  // namespace asb

using namespace asb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int as_embed(const int64_t* tokens, const float* table, int32_t n_vocab, int32_t B,
                        int32_t T, int32_t C, float scale, const int32_t* lens, float* out32,
                        void* out16, int32_t out16_dtype, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(tokens && table && (out32 || out16), AS_ERR_SHAPE, "as_embed: null pointer");
  ASB_REQUIRE(!out16 || out16_dtype == AS_F16 || out16_dtype == AS_BF16, AS_ERR_DTYPE, "as_embed: out16 dtype");
  ASB_CUDA(launch_k(embed_kernel, B * T, 128, 0, ST(stream), tokens, table, n_vocab, B, T, C, scale, lens, out32, out16, out16_dtype));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_layernorm(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                            int32_t C, const float* gamma, const float* beta, float eps,
                            int32_t act, float slope, const int32_t* lens, void* out_a,
                            int32_t out_a_dtype, int64_t out_a_ld, void* out_b,
                            int32_t out_b_dtype, int64_t out_b_ld, void* stream) {
  const long long rows = (long long)B * T;
  if (rows == 0) return AS_OK;
  ASB_REQUIRE(x && gamma && beta && (out_a || out_b), AS_ERR_SHAPE, "as_layernorm: null pointer");
  ASB_REQUIRE(C > 0 && C <= 32 * LN_MAX_PER_LANE, AS_ERR_SHAPE, "as_layernorm: C=%d unsupported", C);
  ASB_REQUIRE(dt_ok(x_dtype), AS_ERR_DTYPE, "as_layernorm: dtype");
  {
    auto ok4 = [](const void* ptr, long long ld, int dt) {
      const int es = dt == AS_F32 ? 4 : 2;
      return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) % (4 * es)) == 0 && ((ld * es) % (4 * es)) == 0);
    };
    if (x_dtype == AS_F32 && (C == 256 || C == 512 || C == 1024) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_ld % 4) == 0 &&
        (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0 &&
        ok4(out_a, out_a_ld, out_a_dtype) && ok4(out_b, out_b_ld, out_b_dtype)) {
      const float* xf = reinterpret_cast<const float*>(x);
#define LNV_LAUNCH(NV_)                                                                                                            \
  ASB_CUDA(launch_k(layernorm_vec4_kernel<NV_>, cdiv(rows, 8), 256, 0, ST(stream), xf, x_ld, rows, T, gamma, beta, eps, act, slope, lens, \
                    out_a, out_a_dtype, out_a_ld, out_b, out_b_dtype, out_b_ld))
      if (C == 256) LNV_LAUNCH(2); else if (C == 512) LNV_LAUNCH(4); else LNV_LAUNCH(8);
#undef LNV_LAUNCH
      ASB_CUDA(cudaGetLastError());
      return AS_OK;
    }
  }
  ASB_CUDA(launch_k(layernorm_kernel, cdiv(rows, 8), 256, 0, ST(stream), x, x_dtype, x_ld, rows, T, C, gamma, beta, eps, act,
                                                         slope, lens, out_a, out_a_dtype, out_a_ld, out_b,
                                                         out_b_dtype, out_b_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_instnorm_stats(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                 int32_t C, const int32_t* lens, float eps, float* stats,
                                 void* stream) {
  if (B * C == 0) return AS_OK;
  ASB_REQUIRE(x && stats && dt_ok(x_dtype), AS_ERR_SHAPE, "as_instnorm_stats: bad argument");
  dim3 grid(cdiv(C, 32), B), block(32, 8);
  ASB_CUDA(launch_k(instnorm_stats_kernel, grid, block, 0, ST(stream), x, x_dtype, x_ld, T, C, lens, eps, stats));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_adain_apply(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                              int32_t C, const float* stats, const float* gb, int64_t gb_ld,
                              float slope, const int32_t* lens, const float* up_w,
                              const float* up_b, void* out, int32_t out_dtype, int64_t out_ld,
                              void* stream) {
  const long long total = (long long)B * T * C * (up_w ? 2 : 1);
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && stats && gb && out, AS_ERR_SHAPE, "as_adain_apply: null pointer");
  ASB_REQUIRE(!up_w || up_b, AS_ERR_SHAPE, "as_adain_apply: up_w without up_b");
  {
    auto row_ok = [](const void* ptr, long long ld, int dt) {
      const int es = dt == AS_F32 ? 4 : 2;
      return (reinterpret_cast<uintptr_t>(ptr) % (4 * es)) == 0 && ((ld * es) % (4 * es)) == 0;
    };
    if ((C % 4) == 0 && row_ok(x, x_ld, x_dtype) && row_ok(out, out_ld, out_dtype)) {
      dim3 grid((unsigned)cdiv(C, 128), (unsigned)cdiv(T, AD_ROWS), (unsigned)B);
      if (up_w) ASB_CUDA(launch_k(adain_apply_vec_kernel<true>, grid, 256, 0, ST(stream), x, x_dtype, x_ld, T, C, stats, gb, gb_ld, slope, lens,
                                                                           up_w, up_b, out, out_dtype, out_ld));
      else ASB_CUDA(launch_k(adain_apply_vec_kernel<false>, grid, 256, 0, ST(stream), x, x_dtype, x_ld, T, C, stats, gb, gb_ld, slope, lens,
                                                                       up_w, up_b, out, out_dtype, out_ld));
      ASB_CUDA(cudaGetLastError());
      return AS_OK;
    }
  }
  ASB_CUDA(launch_k(adain_apply_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, C, stats, gb, gb_ld, slope,
                                                            lens, up_w, up_b, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_adain_norm_apply(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                   int32_t C, const float* gb, int64_t gb_ld, float eps, float slope,
                                   const int32_t* lens, const float* up_w, const float* up_b, void* out,
                                   int32_t out_dtype, int64_t out_ld, float* stats, void* stream) {
  if ((long long)B * T * C == 0) return AS_OK;
  ASB_REQUIRE(x && gb && out && stats && dt_ok(x_dtype), AS_ERR_SHAPE, "as_adain_norm_apply: bad argument");
  ASB_REQUIRE(!up_w || up_b, AS_ERR_SHAPE, "as_adain_norm_apply: up_w without up_b");
  auto row_ok = [](const void* ptr, long long ld, int dt) {
    const int es = dt == AS_F32 ? 4 : 2;
    return (reinterpret_cast<uintptr_t>(ptr) % (4 * es)) == 0 && ((ld * es) % (4 * es)) == 0;
  };
  static const bool no_tma = getenv("ASB_ADAIN_NO_TMA") != nullptr;
  if (!no_tma && x_dtype == AS_F32 && T <= ADT_MAX_ROWS && (C % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (x_ld % 4) == 0 && [&] { const int es = out_dtype == AS_F32 ? 4 : 2;
                              return (reinterpret_cast<uintptr_t>(out) % (4 * es)) == 0 && ((out_ld * es) % (4 * es)) == 0; }() &&
      (reinterpret_cast<uintptr_t>(gb) & 15) == 0 && (gb_ld % 4) == 0 && check_arch() == AS_OK) {
    EncodeTiledFn enc = get_encode_fn();
    static const bool no_ring = getenv("ASB_ADAIN_NO_RING") != nullptr;
    if (enc && !no_ring) {
      // ring of S slabs [nbox * BR][W] fp32 in one persistent CTA per SM (see adain_ring_kernel).  Slab width: 32
      // channels (128-byte rows); 16 channels when a 32-wide slab is too long for a two-stage ring (T > ~860), which
      // takes sequences up to ~1700 frames from 2.8 to 4.0 TB/s.  (Measured and rejected: 16-wide slabs to fill the
      // last round of the static unit schedule better -- 16 x 1024 channels are 3.46 units per SM with W = 32 and 6.92
      // with W = 16 -- the 64-byte rows cost more than the tail: 22.4 vs 17.7 us at 16 x 800 x 1024.)
      const int nbox = (T + 255) / 256;
      const int BR = ((T + nbox - 1) / nbox + 7) / 8 * 8;
      const int nsm = num_sms();
      static const int force_w = getenv("ASB_ADAIN_SLAB") ? atoi(getenv("ASB_ADAIN_SLAB")) : 0;
      const bool two_stages_32 = (size_t)(216 * 1024) / ((size_t)nbox * BR * 128) >= 2;
      const int W = force_w ? force_w : ((!two_stages_32 && (C % 16) == 0) ? 16 : 32);
      const size_t slab_bytes = (size_t)nbox * BR * 4 * W;
      const int S = (int)std::min<size_t>(ADR_MAX_STAGES, (size_t)(216 * 1024) / slab_bytes);
      CUtensorMap tmx;
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)x_ld * 4, (cuuint64_t)x_ld * 4 * T};
      cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)BR, 1};
      cuuint32_t es3[3] = {1, 1, 1};
      if (S >= 2 && enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(x), dims, strides, box, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        const int nslab = cdiv(C, W), units = nslab * B;
        const size_t rsmem = (size_t)S * slab_bytes + 128;
        dim3 rgrid((unsigned)std::min(units, nsm));
        const bool mx = slope >= 0.f && slope <= 1.f;
#define ADR_LAUNCH(UP_, MX_, W_)                                                                                                \
  do {                                                                                                                          \
    ASB_SMEM_OPT_IN(217 * 1024, adain_ring_kernel<UP_, MX_, W_>);                                                               \
    ASB_CUDA(launch_k(adain_ring_kernel<UP_, MX_, W_>, rgrid, ADR_THREADS, rsmem, ST(stream), tmx, T, BR, nbox, C, nslab, units, S, gb, gb_ld, \
                      eps, slope, lens, up_w, up_b, out, out_dtype, out_ld, stats));                                            \
  } while (0)
#define ADR_LAUNCH_W(UP_, MX_) do { if (W == 16) ADR_LAUNCH(UP_, MX_, 16); else ADR_LAUNCH(UP_, MX_, 32); } while (0)
        if (up_w) { if (mx) ADR_LAUNCH_W(true, true); else ADR_LAUNCH_W(true, false); }
        else { if (mx) ADR_LAUNCH_W(false, true); else ADR_LAUNCH_W(false, false); }
#undef ADR_LAUNCH_W
#undef ADR_LAUNCH
        ASB_CUDA(cudaGetLastError());
        return AS_OK;
      }
    }
    if (enc) {
      const int nbox = (T + 255) / 256;
      const int BR = ((T + nbox - 1) / nbox + 7) / 8 * 8;
      CUtensorMap tmx;
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)x_ld * 4, (cuuint64_t)x_ld * 4 * T};
      cuuint32_t box[3] = {32, (cuuint32_t)BR, 1};
      cuuint32_t es3[3] = {1, 1, 1};
      if (enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(x), dims, strides, box, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        const size_t slab_bytes = (size_t)nbox * BR * 128;
        const size_t smem = slab_bytes + 128;
        ASB_SMEM_OPT_IN(225 * 1024, adain_tma_kernel<true>);
        ASB_SMEM_OPT_IN(225 * 1024, adain_tma_kernel<false>);
        dim3 grid(cdiv(C, 32), (unsigned)B);
        if (up_w) ASB_CUDA(launch_k(adain_tma_kernel<true>, grid, 256, smem, ST(stream), tmx, T, BR, nbox, C, gb, gb_ld, eps, slope, lens, up_w, up_b,
                                                                         out, out_dtype, out_ld, stats));
        else ASB_CUDA(launch_k(adain_tma_kernel<false>, grid, 256, smem, ST(stream), tmx, T, BR, nbox, C, gb, gb_ld, eps, slope, lens, up_w, up_b,
                                                                     out, out_dtype, out_ld, stats));
        ASB_CUDA(cudaGetLastError());
        return AS_OK;
      }
    }
  }
  const int TR = ((T + ADF_CS - 1) / ADF_CS + 31) / 32 * 32;     // rows per CTA of the cluster
  if (TR <= ADF_MAX_TR && (C % 4) == 0 && row_ok(x, x_ld, x_dtype) && row_ok(out, out_ld, out_dtype) &&
      (reinterpret_cast<uintptr_t>(gb) & 15) == 0 && (gb_ld % 4) == 0) {
    const size_t smem = (size_t)TR * 32 * sizeof(float);
    dim3 grid(cdiv(C, 32) * ADF_CS, (unsigned)B);
#define ADF_LAUNCH(UP_, XDT_)                                                                                  \
  do {                                                                                                         \
    ASB_SMEM_OPT_IN(ADF_MAX_TR * 32 * 4, adain_fused_kernel<UP_, XDT_>);                                       \
    ASB_CUDA(launch_k(adain_fused_kernel<UP_, XDT_>, grid, 256, smem, ST(stream), x, x_ld, T, TR, C, gb, gb_ld, eps, slope, lens, up_w, \
                                                                   up_b, out, out_dtype, out_ld, stats));       \
  } while (0)
    if (up_w) {
      if (x_dtype == AS_F32) ADF_LAUNCH(true, AS_F32); else if (x_dtype == AS_F16) ADF_LAUNCH(true, AS_F16); else ADF_LAUNCH(true, AS_BF16);
    } else {
      if (x_dtype == AS_F32) ADF_LAUNCH(false, AS_F32); else if (x_dtype == AS_F16) ADF_LAUNCH(false, AS_F16); else ADF_LAUNCH(false, AS_BF16);
    }
#undef ADF_LAUNCH
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  // long sequences: statistics pass + streaming apply pass
  int rc = as_instnorm_stats(x, x_dtype, x_ld, B, T, C, lens, eps, stats, stream);
  if (rc != AS_OK) return rc;
  return as_adain_apply(x, x_dtype, x_ld, B, T, C, stats, gb, gb_ld, slope, lens, up_w, up_b, out, out_dtype, out_ld, stream);
}

extern "C" int as_repeat_rows(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                              int32_t C, int32_t rep, const int32_t* lens, void* out,
                              int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = (long long)B * T * rep * C;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && out && rep >= 1, AS_ERR_SHAPE, "as_repeat_rows: bad argument");
  ASB_CUDA(launch_k(repeat_rows_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, C, rep, lens, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_expand_tokens(const float* x, const int32_t* token_of_frame, float* out, int32_t B, int32_t C,
                                int32_t Tx, int32_t Ty, void* stream) {
  const long long total = (long long)B * C * Ty;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && token_of_frame && out && Tx > 0, AS_ERR_SHAPE, "as_expand_tokens: bad argument");
  int rc = check_arch();
  if (rc != AS_OK) return rc;
  ASB_CUDA(launch_k(expand_tokens_kernel, ew_grid(total), 256, 0, ST(stream), x, token_of_frame, out, (long long)B * C, C, Tx, Ty));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_length_regulate(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t Tt,
                                  int32_t C, const int32_t* dur, const int32_t* lens_t, int32_t rep,
                                  int32_t To, void* out, int32_t out_dtype, int64_t out_ld,
                                  int32_t* out_lens, void* stream) {
  if (B == 0) return AS_OK;
  ASB_REQUIRE(x && dur && out && rep >= 1 && Tt > 0 && To >= 0, AS_ERR_SHAPE, "as_length_regulate: bad argument");
  const size_t smem = (size_t)(Tt + 1) * sizeof(int);
  ASB_REQUIRE(smem <= 48 * 1024, AS_ERR_SHAPE, "as_length_regulate: Tt=%d too large", Tt);
  auto al16 = [](const void* ptr, long long ld, int dt) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((ld * (dt == AS_F32 ? 4 : 2)) & 15) == 0;
  };
  int vec = 0;
  if (al16(x, x_ld, x_dtype) && al16(out, out_ld, out_dtype)) {
    if (x_dtype == out_dtype && (C * (x_dtype == AS_F32 ? 4 : 2)) % 16 == 0) vec = 1;
    else if (x_dtype == AS_F32 && out_dtype != AS_F32 && C % 8 == 0) vec = 2;
  }
  dim3 grid((unsigned)((To + LR_ROWS - 1) / LR_ROWS > 0 ? (To + LR_ROWS - 1) / LR_ROWS : 1), (unsigned)B);
  ASB_CUDA(launch_k(length_regulate_kernel, grid, 256, smem, ST(stream), x, x_dtype, x_ld, Tt, C, dur, lens_t, rep, To, out,
                                                         out_dtype, out_ld, out_lens, vec));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_conv_small(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                             int32_t F, int32_t Cin, const float* w, const float* bias,
                             int32_t ntaps, const int32_t* tap_dt, const int32_t* tap_df,
                             int32_t Cout, const int32_t* lens, void* y_raw, int32_t y_raw_dtype,
                             int64_t y_raw_ld, void* y_act, int32_t y_act_dtype, int64_t y_act_ld,
                             int32_t act, float slope, void* stream) {
  const long long total = (long long)B * T * F * Cout;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && w && (y_raw || y_act) && tap_dt && tap_df, AS_ERR_SHAPE, "as_conv_small: null pointer");
  ASB_REQUIRE(ntaps >= 1 && ntaps <= CS_MAX_TAPS && Cin >= 1 && Cin <= 16, AS_ERR_SHAPE,
              "as_conv_small: ntaps=%d Cin=%d unsupported", ntaps, Cin);
  SmallTaps taps;
  for (int j = 0; j < CS_MAX_TAPS; ++j) { taps.dt[j] = j < ntaps ? tap_dt[j] : 0; taps.df[j] = j < ntaps ? tap_df[j] : 0; }
  const size_t wsmem = ((size_t)ntaps * Cin * Cout + Cout) * sizeof(float);
  if ((Cout % 8) == 0 && wsmem <= 40 * 1024 && total < (1ll << 31)) {
    auto al16 = [](const void* ptr, long long ld, int dt) {
      return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((ld * (dt == AS_F32 ? 4 : 2)) & 15) == 0;
    };
    if ((F % 4) == 0)
      ASB_CUDA(launch_k(conv_small_vec_kernel<4>, ew_grid(total / 32), 256, wsmem, ST(stream),
          x, x_dtype, x_ld, B, T, F, Cin, w, bias, ntaps, taps, Cout, lens, y_raw, y_raw_dtype, y_raw_ld, y_act,
          y_act_dtype, y_act_ld, act, slope, al16(y_raw, y_raw_ld, y_raw_dtype) ? 1 : 0, al16(y_act, y_act_ld, y_act_dtype) ? 1 : 0));
    else
      ASB_CUDA(launch_k(conv_small_vec_kernel<1>, ew_grid(total / 8), 256, wsmem, ST(stream),
          x, x_dtype, x_ld, B, T, F, Cin, w, bias, ntaps, taps, Cout, lens, y_raw, y_raw_dtype, y_raw_ld, y_act,
          y_act_dtype, y_act_ld, act, slope, al16(y_raw, y_raw_ld, y_raw_dtype) ? 1 : 0, al16(y_act, y_act_ld, y_act_dtype) ? 1 : 0));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  ASB_CUDA(launch_k(conv_small_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, Cin, w, bias, ntaps, taps,
                                                           Cout, lens, y_raw, y_raw_dtype, y_raw_ld, y_act,
                                                           y_act_dtype, y_act_ld, act, slope));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_dwconv(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T, int32_t F,
                         int32_t C, int32_t glu, const float* w, const float* bias, int32_t kt,
                         int32_t kf, int32_t st, int32_t sf, int32_t pt, int32_t pf, int32_t To,
                         int32_t Fo, const int32_t* lens_in, const int32_t* lens_out, int32_t act,
                         float slope, void* out, int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = (long long)B * To * Fo * C;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && w && out && kt >= 1 && kf >= 1 && st >= 1 && sf >= 1, AS_ERR_SHAPE, "as_dwconv: bad argument");
  if (F == 1 && Fo == 1 && kf == 1 && sf == 1 && st == 1 && pf == 0 && To == T && kt <= DWT_MAXK && kt >= 5) {
    dim3 grid(cdiv(C, DWT_CH), cdiv(T, DWT_ROWS), (unsigned)B);
    const size_t smem = (size_t)(DWT_ROWS + kt - 1) * DWT_CH * sizeof(float);
    ASB_CUDA(launch_k(dwconv_time_tiled_kernel, grid, 256, smem, ST(stream), x, x_dtype, x_ld, T, C, glu, w, bias, kt, pt, lens_in, lens_out,
                                                            act, slope, out, out_dtype, out_ld));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  if (!glu && vec8_ok(x, x_dtype, x_ld, out, out_dtype, out_ld, C, total) && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    ASB_CUDA(launch_k(dwconv_vec8_kernel, ew_grid(total / 8), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, w, bias, kt, kf, st, sf,
                      pt, pf, To, Fo, lens_in, lens_out, act, slope, out, out_dtype, out_ld));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  ASB_CUDA(launch_k(dwconv_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, glu, w, bias, kt, kf, st, sf,
                                                       pt, pf, To, Fo, lens_in, lens_out, act, slope, out,
                                                       out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_avgpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T, int32_t F,
                          int32_t C, int32_t pt, int32_t pf, void* out, int32_t out_dtype,
                          int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && pt >= 1 && pf >= 1, AS_ERR_SHAPE, "as_avgpool: bad argument");
  const int To = (T + pt - 1) / pt, Fo = F / pf;
  const long long total = (long long)B * To * Fo * C;
  if (total == 0) return AS_OK;
  if (vec8_ok(x, x_dtype, x_ld, out, out_dtype, out_ld, C, total)) {
    ASB_CUDA(launch_k(avgpool_vec8_kernel, ew_grid(total / 8), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, pt, pf, To, Fo, out, out_dtype,
                      out_ld));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  ASB_CUDA(launch_k(avgpool_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, pt, pf, To, Fo, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_affine_act_maxpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                     int32_t F, int32_t C, const float* scale, const float* shift,
                                     float slope, int32_t pf, void* out, int32_t out_dtype,
                                     int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && scale && shift && pf >= 1, AS_ERR_SHAPE, "as_affine_act_maxpool: bad argument");
  const int Fo = F / pf;
  const long long total = (long long)B * T * Fo * C;
  if (total == 0) return AS_OK;
  if (vec8_ok(x, x_dtype, x_ld, out, out_dtype, out_ld, C, total)) {
    ASB_CUDA(launch_k(affine_act_maxpool_vec8_kernel, ew_grid(total / 8), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, scale, shift,
                      slope, pf, Fo, out, out_dtype, out_ld));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  ASB_CUDA(launch_k(affine_act_maxpool_kernel, ew_grid(total), 256, 0, ST(stream), x, x_dtype, x_ld, B, T, F, C, scale, shift,
                                                                   slope, pf, Fo, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_global_avgpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                 int32_t F, int32_t C, int32_t t_stride, float slope, void* out,
                                 int32_t out_dtype, int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && t_stride >= 1 && T >= 1 && F >= 1, AS_ERR_SHAPE, "as_global_avgpool: bad argument");
  if (B * C == 0) return AS_OK;
  dim3 grid(cdiv(C, 128), B);
  ASB_CUDA(launch_k(global_avgpool_kernel, grid, 128, 0, ST(stream), x, x_dtype, x_ld, T, F, C, t_stride, slope, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_lstm_onestep(const float* xproj, int64_t xproj_ld, int64_t rows, int32_t H, void* out,
                               int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = rows * 2 * H;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(xproj && out, AS_ERR_SHAPE, "as_lstm_onestep: null pointer");
  ASB_CUDA(launch_k(lstm_onestep_kernel, ew_grid(total), 256, 0, ST(stream), xproj, xproj_ld, rows, H, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_log_norm(const float* mel, int32_t B, int32_t n_mels, int32_t T, float* out, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(mel && out, AS_ERR_SHAPE, "as_log_norm: null pointer");
  ASB_CUDA(launch_k(log_norm_kernel, cdiv((long long)B * T, 128), 128, 0, ST(stream), mel, B, n_mels, T, out));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_transpose_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype,
                                 int32_t B, int32_t C, int32_t T, int64_t cl_ld,
                                 int32_t to_channels_last, const float* sub, const float* mul,
                                 const int32_t* lens, void* stream) {
  if (B * C * T == 0) return AS_OK;
  ASB_REQUIRE(src && dst && dt_ok(src_dtype) && dt_ok(dst_dtype), AS_ERR_SHAPE, "as_transpose_cast: bad argument");
  ASB_REQUIRE((sub == nullptr) == (mul == nullptr), AS_ERR_SHAPE, "as_transpose_cast: sub/mul must come together");
  dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
  ASB_CUDA(launch_k(transpose_cast_kernel, grid, block, 0, ST(stream), src, src_dtype, dst, dst_dtype, C, T, cl_ld,
                                                       to_channels_last, sub, mul, lens));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
