// Attention kernels (fp32 on CUDA cores; attention is <3 % of the path's FLOPs).
//   relpos_attention    — windowed relative-position MHA of RelTransformerEnc.py:138-169, fused:
//                         QK^T + rel-key bias, length mask, softmax, PV + rel-value term.
//   conformer_attention — Transformer-XL attention with the reference's view-based relative shift
//                         (Utils/EMA/conformer/conformer/attention.py:72-113).
// Neither materialises the [B,H,T,T] score tensor in HBM (the reference does, plus pad/view skew
// copies); scores live in shared memory per query block.
#include <type_traits>

#include "common.cuh"

namespace asb {

// ---------------------------------------------------------------------------------------------
// relpos attention.  CTA = 8 warps, each warp owns QW = 4 consecutive queries of one (b, h);
// K / V tiles of 32 keys are staged in shared memory and shared by the CTA's 32 queries (float4 inner loops:
// one 16-byte key load + QW broadcast query loads per 4 x QW FMAs).
// ---------------------------------------------------------------------------------------------
constexpr int RA_WARPS = 8;
constexpr int RA_QW = 4;
constexpr int RA_QT = RA_WARPS * RA_QW;  // queries per CTA
constexpr int RA_KT = 32;                // keys per tile
constexpr int RA_MAXW = 9;               // 2*window+1 <= 9

// 3 CTAs per SM (<= 85 registers): the bench shape launches 320 CTAs; at 2 per SM the last 24 made a second wave
template <int D>
__global__ void __launch_bounds__(RA_WARPS * 32, 3)
relpos_attention_kernel(const float* __restrict__ qkv, long long ld, const float* __restrict__ relk,
                        const float* __restrict__ relv, int window, int T, int H,
                        const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(16) float sm[];
  constexpr int KP = D + 4;                 // padded pitch: 16-byte aligned, conflict-free float4 row reads
  float* kv = sm;                           // [RA_KT][KP]
  float* qs = kv + RA_KT * KP;              // [RA_QT][D]
  float* qe = qs + RA_QT * D;               // [RA_QT][RA_MAXW] q . E_k[r]
  float* rks = qe + RA_QT * RA_MAXW;        // [RA_MAXW][D] relative-key embeddings  (staged: read from global they
  float* rvs = rks + RA_MAXW * D;           // [RA_MAXW][D] relative-value embeddings  cost ~290 dependent loads per warp)
  float* sc = rvs + RA_MAXW * D;            // [RA_QT][Tpad] scores / probabilities
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * RA_QT;
  const int len = lens ? min(lens[b], T) : T;
  const int Tpad = (T + 31) & ~31;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrel = 2 * window + 1;
  const float scale = rsqrtf((float)D);
  const float* base = qkv + (long long)b * T * ld;
  const int HD = H * D;

  if (q0 >= len) {
    // whole query block is padding: write zeros
    for (int i = tid; i < RA_QT * D; i += blockDim.x) {
      int qi = q0 + i / D, d = i % D;
      if (qi < T) stany(out, ((long long)b * T + qi) * out_ld + h * D + d, 0.f, odt);
    }
    return;
  }

  // stage the relative embeddings and the CTA's queries
  for (int i = tid; i < RA_MAXW * D; i += blockDim.x) {
    const bool in = i < nrel * D;
    rks[i] = in ? relk[i] : 0.f;
    rvs[i] = in ? relv[i] : 0.f;
  }
  for (int i = tid; i < RA_QT * D; i += blockDim.x) {
    int qi = q0 + i / D, d = i % D;
    qs[i] = qi < T ? base[(long long)qi * ld + h * D + d] : 0.f;
  }
  __syncthreads();
  // q . E_k[r]
  for (int i = warp; i < RA_QT * RA_MAXW; i += RA_WARPS) {
    int ql = i / RA_MAXW, r = i % RA_MAXW;
    float s = 0.f;
    if (r < nrel)
      for (int d = lane; d < D; d += 32) s += qs[ql * D + d] * rks[r * D + d];
    s = warp_sum(s);
    if (lane == 0) qe[i] = s;
  }
  __syncthreads();

  // K / V tile staging: 16-byte cp.async with zero fill for keys >= len, every copy of the tile in flight at once
  // (with 4-byte loads through registers the 16 dependent loads per thread and tile set the kernel's time: ~90 us
  // for 0.15 GFLOP at T = 150); scalar loop when the rows are not 16-byte aligned
  const bool vec = (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0 && (HD & 3) == 0;
  auto stage_tile = [&](int kt, int off) {
    if (vec) {
      const uint32_t kv_u = smem_u32(kv);
      for (int i = tid; i < RA_KT * (D / 4); i += blockDim.x) {
        const int r = i / (D / 4), c = (i - r * (D / 4)) * 4;
        const int j = kt * RA_KT + r;
        const bool ok = j < len;
        const float* src = ok ? base + (long long)j * ld + off + h * D + c : base;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(kv_u + (uint32_t)(r * KP + c) * 4u), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    } else {
      for (int i = tid; i < RA_KT * D; i += blockDim.x) {
        int j = kt * RA_KT + i / D, d = i % D;
        kv[(i / D) * KP + d] = j < len ? base[(long long)j * ld + off + h * D + d] : 0.f;
      }
    }
    __syncthreads();
  };

  // ---- scores ----
  const int ntiles = (len + RA_KT - 1) / RA_KT;
  for (int kt = 0; kt < ntiles; ++kt) {
    stage_tile(kt, HD);
    const int j = kt * RA_KT + lane;
    float s[RA_QW];
#pragma unroll
    for (int u = 0; u < RA_QW; ++u) s[u] = 0.f;
    const float* kr = kv + lane * KP;
#pragma unroll 4
    for (int d = 0; d < D; d += 4) {
      const float4 kd = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
      for (int u = 0; u < RA_QW; ++u) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + (warp * RA_QW + u) * D + d);
        s[u] += q4.x * kd.x + q4.y * kd.y + q4.z * kd.z + q4.w * kd.w;
      }
    }
#pragma unroll
    for (int u = 0; u < RA_QW; ++u) {
      const int ql = warp * RA_QW + u, qi = q0 + ql;
      float v = s[u];
      const int rel = j - qi;
      if (rel >= -window && rel <= window) v += qe[ql * RA_MAXW + rel + window];
      v *= scale;
      if (j < Tpad) sc[ql * Tpad + j] = (j < len) ? v : -INFINITY;
    }
    __syncthreads();
  }

  // ---- softmax over keys < len (masked keys contribute exp(-1e4 - max) == 0 in fp32) ----
#pragma unroll
  for (int u = 0; u < RA_QW; ++u) {
    const int ql = warp * RA_QW + u;
    float* row = sc + ql * Tpad;
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < len; j += 32) { float e = expf(row[j] - m); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < len; j += 32) row[j] *= inv;
  }
  __syncthreads();

  // ---- P . V (+ relative values) ----
  constexpr int DPL = D / 32;  // output dims per lane
  float acc[RA_QW][DPL];
#pragma unroll
  for (int u = 0; u < RA_QW; ++u)
#pragma unroll
    for (int c = 0; c < DPL; ++c) acc[u][c] = 0.f;
  for (int kt = 0; kt < ntiles; ++kt) {
    stage_tile(kt, 2 * HD);
    const int jn = min(RA_KT, len - kt * RA_KT);
    for (int jj = 0; jj < jn; ++jj) {
      float p[RA_QW];
#pragma unroll
      for (int u = 0; u < RA_QW; ++u) p[u] = sc[(warp * RA_QW + u) * Tpad + kt * RA_KT + jj];
#pragma unroll
      for (int c = 0; c < DPL; ++c) {
        const float vv = kv[jj * KP + lane + 32 * c];
#pragma unroll
        for (int u = 0; u < RA_QW; ++u) acc[u][c] += p[u] * vv;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < RA_QW; ++u) {
    const int ql = warp * RA_QW + u, qi = q0 + ql;
    if (qi >= T) continue;
    const bool valid = qi < len;
    if (valid) {
      for (int r = -window; r <= window; ++r) {
        const int j = qi + r;
        if (j < 0 || j >= len) continue;
        const float p = sc[ql * Tpad + j];
#pragma unroll
        for (int c = 0; c < DPL; ++c) acc[u][c] += p * rvs[(r + window) * D + lane + 32 * c];
      }
    }
#pragma unroll
    for (int c = 0; c < DPL; ++c)
      stany(out, ((long long)b * T + qi) * out_ld + h * D + lane + 32 * c, valid ? acc[u][c] : 0.f, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// conformer attention: one warp per (b, h, query); K / pos / V rows are read straight from
// global memory (L1/L2 resident: T*H*D*4 bytes per tensor).
// ---------------------------------------------------------------------------------------------
constexpr int CA_WARPS = 4;

template <int D>
__global__ void __launch_bounds__(CA_WARPS * 32)
conformer_attention_kernel(const float* __restrict__ q, const float* __restrict__ k,
                           const float* __restrict__ v, long long ld, const float* __restrict__ pos,
                           const float* __restrict__ ub, const float* __restrict__ vb, int T, int H,
                           const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int a = blockIdx.x * CA_WARPS + warp;  // query
  const int len = lens ? min(lens[b], T) : T;
  const int Tpad = (T + 31) & ~31;
  float* qu = sm + warp * (3 * D + Tpad);   // q_a + u
  float* qv0 = qu + D;                      // q_a + v
  float* qv1 = qv0 + D;                     // q_{a+1} + v
  float* sc = qv1 + D;                      // [Tpad]
  if (a >= T) return;
  const int HD = H * D;
  if (a >= len) {
    for (int d = lane; d < D; d += 32) stany(out, ((long long)b * T + a) * out_ld + h * D + d, 0.f, odt);
    return;
  }
  const float* qb = q + (long long)b * T * ld + h * D;
  const float* kb = k + (long long)b * T * ld + h * D;
  const float* vbase = v + (long long)b * T * ld + h * D;
  for (int d = lane; d < D; d += 32) {
    const float qa = qb[(long long)a * ld + d];
    const float qn = (a + 1 < len) ? qb[(long long)(a + 1) * ld + d] : 0.f;
    qu[d] = qa + ub[h * D + d];
    qv0[d] = qa + vb[h * D + d];
    qv1[d] = qn + vb[h * D + d];
  }
  __syncwarp();
  const float inv_sqrt = rsqrtf((float)HD);
  // the shift is defined on this item's own [len x len] score matrix
  const int L = len;
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const float* kr = kb + (long long)j * ld;
    float s = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(kr + d);
      s += qu[d] * kk.x + qu[d + 1] * kk.y + qu[d + 2] * kk.z + qu[d + 3] * kk.w;
    }
    // relative shift (attention.py:105-113): flat index of out[a][j] in the padded tensor
    const long long f = (long long)(a + 1) * L + j;
    const int i2 = (int)(f / (L + 1)), jj = (int)(f % (L + 1));
    if (jj != 0) {
      const float* pr = pos + (long long)(jj - 1) * HD + h * D;
      const float* qq = (i2 == a) ? qv0 : qv1;
      float ps = 0.f;
#pragma unroll 4
      for (int d = 0; d < D; d += 4) {
        const float4 pp = *reinterpret_cast<const float4*>(pr + d);
        ps += qq[d] * pp.x + qq[d + 1] * pp.y + qq[d + 2] * pp.z + qq[d + 3] * pp.w;
      }
      s += ps;
    }
    s *= inv_sqrt;
    sc[j] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) { float e = expf(sc[j] - m); sc[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __syncwarp();
  constexpr int DPL = D / 32;
  float acc[DPL];
#pragma unroll
  for (int c = 0; c < DPL; ++c) acc[c] = 0.f;
  for (int j = 0; j < L; ++j) {
    const float p = sc[j];
#pragma unroll
    for (int c = 0; c < DPL; ++c) acc[c] += p * vbase[(long long)j * ld + lane + 32 * c];
  }
#pragma unroll
  for (int c = 0; c < DPL; ++c)
    stany(out, ((long long)b * T + a) * out_ld + h * D + lane + 32 * c, acc[c] * inv, odt);
}

// ---------------------------------------------------------------------------------------------
// Tiled conformer attention (the per-query kernel above re-reads K / pos / V from L1/L2 for every
// query with one row per lane: 545 us for B=16, T=240 under ncu).  One CTA per (b, h, 32 queries):
// K, pos and V stream once through a shared 32-row tile; content scores S = (Q+u) K^T and position
// scores M = (Q+v) P^T are two small GEMMs (lane <-> key, 4 queries per warp), the Transformer-XL
// relative shift (attention.py:105-113) is a gather from M: out[a][j] = M[i2][jj-1] with
// (i2, jj) = divmod((a+1)*L + j, L+1), zero when jj == 0.
// ---------------------------------------------------------------------------------------------
constexpr int CT_QT = 32;       // queries per CTA
constexpr int CT_WARPS = 8;     // 4 queries per warp

template <int D>
__global__ void __launch_bounds__(CT_WARPS * 32)
conformer_attention_tiled_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                 const float* __restrict__ v, long long ld, const float* __restrict__ pos,
                                 const float* __restrict__ ub, const float* __restrict__ vb, int T, int H,
                                 const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int KS = D + 4;     // padded row stride of the key tile (float4-aligned, conflict-free)
  constexpr int QPW = CT_QT / CT_WARPS;
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, a0 = blockIdx.x * CT_QT;
  const int len = lens ? min(lens[b], T) : T;
  const int Tpad = (T + 31) & ~31;
  float* Qu = sm;                           // [QT][D]      q + u
  float* Qv = Qu + CT_QT * D;               // [QT+1][D]    q + v (rows a0 .. a0+QT)
  float* KT = Qv + (CT_QT + 1) * D;         // [32][KS]
  float* S = KT + 32 * KS;                  // [QT][Tpad]
  float* Mx = S + CT_QT * Tpad;             // [QT+1][Tpad]
  const int HD = H * D;
  if (a0 >= len) {
    for (int i = threadIdx.x; i < CT_QT * D; i += blockDim.x) {
      const int r = a0 + i / D;
      if (r < T) stany(out, ((long long)b * T + r) * out_ld + h * D + (i % D), 0.f, odt);
    }
    return;
  }
  const float* qb = q + (long long)b * T * ld + h * D;
  const float* kb = k + (long long)b * T * ld + h * D;
  const float* vbase = v + (long long)b * T * ld + h * D;
  for (int i = threadIdx.x; i < (CT_QT + 1) * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D, a = a0 + r;
    const float qa = a < len ? qb[(long long)a * ld + d] : 0.f;
    if (r < CT_QT) Qu[i] = qa + ub[h * D + d];
    Qv[i] = qa + vb[h * D + d];
  }
  const int L = len;
  auto load_tile = [&](const float* src, long long stride, int j0) {
    for (int i = threadIdx.x; i < 32 * (D / 4); i += blockDim.x) {
      const int r = i / (D / 4), c = i - r * (D / 4);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < L) t = *reinterpret_cast<const float4*>(src + (long long)(j0 + r) * stride + 4 * c);
      *reinterpret_cast<float4*>(KT + r * KS + 4 * c) = t;
    }
  };
  // NR rows of `Qm` (this warp's share) dotted with the 32 keys of the tile; NR is a compile-time constant so that the
  // extra row only warp 7 needs costs the other warps nothing
  auto dot_rows = [&](auto nr_tag, const float* Qm, int r0, float* dst, int j0) {
    constexpr int NR = decltype(nr_tag)::value;
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] = 0.f;
    const float* kr = KT + lane * KS;
#pragma unroll 4
    for (int d = 0; d < D; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const float4 qq = *reinterpret_cast<const float4*>(Qm + (r0 + i) * D + d);
        acc[i] = fmaf(qq.x, kk.x, fmaf(qq.y, kk.y, fmaf(qq.z, kk.z, fmaf(qq.w, kk.w, acc[i]))));
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) dst[(r0 + i) * Tpad + j0 + lane] = acc[i];
  };
  for (int j0 = 0; j0 < L; j0 += 32) {
    __syncthreads();
    load_tile(kb, ld, j0);
    __syncthreads();
    dot_rows(std::integral_constant<int, QPW>{}, Qu, warp * QPW, S, j0);
    __syncthreads();
    load_tile(pos + h * D, HD, j0);
    __syncthreads();
    // warp 7 also takes the extra row QT (query a0+QT, needed by the shift of the tile's last query)
    if (warp == CT_WARPS - 1) dot_rows(std::integral_constant<int, QPW + 1>{}, Qv, warp * QPW, Mx, j0);
    else dot_rows(std::integral_constant<int, QPW>{}, Qv, warp * QPW, Mx, j0);
  }
  __syncthreads();
  const float inv_sqrt = rsqrtf((float)HD);
  float inv_sum[QPW];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    const int r = warp * QPW + i, a = a0 + r;
    inv_sum[i] = 0.f;
    if (a >= len) continue;   // warp-uniform
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      const unsigned f = (unsigned)(a + 1) * (unsigned)L + (unsigned)j;      // < 2^32 for T < 65535: 32-bit divide
      const int i2 = (int)(f / (unsigned)(L + 1)), jj = (int)(f - (unsigned)i2 * (unsigned)(L + 1));
      float sc = S[r * Tpad + j];
      if (jj != 0) sc += Mx[(i2 - a0) * Tpad + jj - 1];
      sc *= inv_sqrt;
      S[r * Tpad + j] = sc;
      m = fmaxf(m, sc);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = expf(S[r * Tpad + j] - m); S[r * Tpad + j] = e; sum += e; }
    inv_sum[i] = 1.f / warp_sum(sum);
  }
  constexpr int DPL = D / 32;
  float acc[QPW][DPL];
#pragma unroll
  for (int i = 0; i < QPW; ++i)
#pragma unroll
    for (int c = 0; c < DPL; ++c) acc[i][c] = 0.f;
  for (int j0 = 0; j0 < L; j0 += 32) {
    __syncthreads();
    load_tile(vbase, ld, j0);
    __syncthreads();
    const int nk = min(32, L - j0);
    for (int jj = 0; jj < nk; ++jj) {
      float vv[DPL];
#pragma unroll
      for (int c = 0; c < DPL; ++c) vv[c] = KT[jj * KS + lane + 32 * c];
#pragma unroll
      for (int i = 0; i < QPW; ++i) {
        const float p = S[(warp * QPW + i) * Tpad + j0 + jj];
#pragma unroll
        for (int c = 0; c < DPL; ++c) acc[i][c] += p * vv[c];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    const int a = a0 + warp * QPW + i;
    if (a >= T) continue;
#pragma unroll
    for (int c = 0; c < DPL; ++c)
      stany(out, ((long long)b * T + a) * out_ld + h * D + lane + 32 * c, a < len ? acc[i][c] * inv_sum[i] : 0.f, odt);
  }
}

}  // namespace asb

namespace asb {
int relpos_attention_mma_launch(const float* qkv, long long ld, const float* relk, const float* relv, int window, int B,
                                int T, int H, const int* lens, void* out, int odt, long long out_ld, cudaStream_t st);
bool conformer_mma_eligible(int T);
int conformer_attention_mma_launch(const float* q, const float* k, const float* v, long long ld, const float* pos,
                                   const float* ub, const float* vb, int B, int T, int H, const int* lens, void* out,
                                   int odt, long long out_ld, cudaStream_t st);
}

using namespace asb;

extern "C" int as_relpos_attention(const float* qkv, int64_t qkv_ld, const float* emb_rel_k,
                                   const float* emb_rel_v, int32_t window, int32_t B, int32_t T,
                                   int32_t H, int32_t D, const int32_t* lens, void* out,
                                   int32_t out_dtype, int64_t out_ld, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(qkv && emb_rel_k && emb_rel_v && out, AS_ERR_SHAPE, "as_relpos_attention: null pointer");
  ASB_REQUIRE(D == 128, AS_ERR_SHAPE, "as_relpos_attention: head dim %d unsupported (128 only)", D);
  ASB_REQUIRE(window >= 0 && 2 * window + 1 <= RA_MAXW, AS_ERR_SHAPE, "as_relpos_attention: window");
  // tensor-core path (mma.sync, fp16 operands / fp32 accumulate and softmax): 16-byte aligned rows, fp32 or
  // 16-bit output; ASB_ATTN_FP32=1 keeps the all-fp32 CUDA-core kernel (numerics reference)
  static const bool fp32_only = getenv("ASB_ATTN_FP32") != nullptr;
  if (!fp32_only && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (qkv_ld % 4) == 0 && (out_ld % 2) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 7) == 0 && (reinterpret_cast<uintptr_t>(emb_rel_k) & 15) == 0)
    return relpos_attention_mma_launch(qkv, qkv_ld, emb_rel_k, emb_rel_v, window, B, T, H, lens, out, out_dtype, out_ld,
                                       reinterpret_cast<cudaStream_t>(stream));
  const int Tpad = (T + 31) & ~31;
  const size_t smem = sizeof(float) * ((size_t)RA_KT * (D + 4) + RA_QT * D + RA_QT * RA_MAXW + 2 * RA_MAXW * D + (size_t)RA_QT * Tpad);
  ASB_REQUIRE(smem <= 200 * 1024, AS_ERR_SHAPE, "as_relpos_attention: T=%d too long for the score buffer", T);
  ASB_SMEM_OPT_IN(200 * 1024, relpos_attention_kernel<128>);
  dim3 grid((T + RA_QT - 1) / RA_QT, H, B);
  ASB_CUDA(launch_k(relpos_attention_kernel<128>, grid, RA_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream), 
      qkv, qkv_ld, emb_rel_k, emb_rel_v, window, T, H, lens, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_conformer_attention(const float* q, const float* k, const float* v, int64_t qkv_ld,
                                      const float* pos, const float* u_bias, const float* v_bias,
                                      int32_t B, int32_t T, int32_t H, int32_t D,
                                      const int32_t* lens, void* out, int32_t out_dtype,
                                      int64_t out_ld, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(q && k && v && pos && u_bias && v_bias && out, AS_ERR_SHAPE, "as_conformer_attention: null pointer");
  ASB_REQUIRE(D == 64, AS_ERR_SHAPE, "as_conformer_attention: head dim %d unsupported (64 only)", D);
  ASB_REQUIRE((qkv_ld % 4) == 0 && ((H * D) % 4) == 0, AS_ERR_ALIGN, "as_conformer_attention: ld must be a multiple of 4");
  {
    // tensor-core path (mma.sync, fp16 operands, fp32 accumulate / softmax); ASB_ATTN_FP32=1 keeps the fp32 kernels
    static const bool fp32_only = getenv("ASB_ATTN_FP32") != nullptr;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!fp32_only && conformer_mma_eligible(T) && al16(q) && al16(k) && al16(v) && al16(pos) && al16(u_bias) && al16(v_bias) &&
        (out_ld % 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
      return conformer_attention_mma_launch(q, k, v, qkv_ld, pos, u_bias, v_bias, B, T, H, lens, out, out_dtype, out_ld,
                                            reinterpret_cast<cudaStream_t>(stream));
  }
  const int Tpad = (T + 31) & ~31;
  {
    const size_t smem_t = sizeof(float) * ((size_t)(2 * CT_QT + 1) * D + 32 * (D + 4) + (size_t)(2 * CT_QT + 1) * Tpad);
    if (smem_t <= 200 * 1024) {
      ASB_SMEM_OPT_IN(200 * 1024, conformer_attention_tiled_kernel<64>);
      dim3 grid((T + CT_QT - 1) / CT_QT, H, B);
      ASB_CUDA(launch_k(conformer_attention_tiled_kernel<64>, grid, CT_WARPS * 32, smem_t, reinterpret_cast<cudaStream_t>(stream), 
          q, k, v, qkv_ld, pos, u_bias, v_bias, T, H, lens, out, out_dtype, out_ld));
      ASB_CUDA(cudaGetLastError());
      return AS_OK;
    }
  }
  const size_t smem = sizeof(float) * (size_t)CA_WARPS * (3 * D + Tpad);
  ASB_REQUIRE(smem <= 200 * 1024, AS_ERR_SHAPE, "as_conformer_attention: T=%d too long", T);
  ASB_SMEM_OPT_IN(200 * 1024, conformer_attention_kernel<64>);
  dim3 grid((T + CA_WARPS - 1) / CA_WARPS, H, B);
  ASB_CUDA(launch_k(conformer_attention_kernel<64>, grid, CA_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream), 
      q, k, v, qkv_ld, pos, u_bias, v_bias, T, H, lens, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
