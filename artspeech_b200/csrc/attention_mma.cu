// Relative-position attention on tensor cores (RelTransformerEnc.py:138-169): QK^T, the windowed rel-key term, the
// softmax-weighted V sum and the rel-value term in one flash-style pass, mma.sync m16n8k16 (fp16 operands, fp32
// accumulation, fp32 online softmax).  The [B,H,T,T] score tensor exists only as register fragments.
//
//   scores[i,j] = (q_i . k_j + [|j-i| <= w] q_i . Ek[j-i+w]) / sqrt(D),  keys j >= len masked
//   out_i       = sum_j p_ij v_j + sum_{|r| <= w} p_{i,i+r} Ev[r+w]
//
// CTA = 4 warps = 64 queries of one (item, head); warp = 16 queries.  K / V stream through shared memory in tiles of
// 64 keys, converted fp32 -> fp16 while they are staged (qkv is the fp32 output of the fused q|k|v projection).
// The rel-key term is the 9-wide side GEMM Q . Ek^T (one extra pair of n8 tiles, computed once per CTA and parked in
// shared memory); the scores inside the +-w band are parked as well, so the rel-value term is 9 FMAs per output
// element after the loop.  At the path's sizes (T <= 1000, D = 128) the kernel is latency-bound: its cost is the
// staging of K / V, not the 2 * 2 * T^2 * D flops -- the CUDA-core version it replaces spent 40 M warp-instructions
// per launch on exactly those flops (profiles/r01_ncu_full_relpos_attention_v49.txt).
#include "common.cuh"

namespace asb {

constexpr int RM_WARPS = 4;
constexpr int RM_QT = 64;        // queries per CTA
constexpr int RM_KT = 64;        // keys per tile
constexpr int RM_NREL = 9;       // 2 * window + 1 <= 9

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// rows x D fp32 (row stride ld) -> fp16 shared tile with pitch P halfs; rows >= nvalid are zero-filled
template <int D, int P>
__device__ __forceinline__ void stage_f16(__half* dst, const float* src, long long ld, int row0, int nrows, int nvalid, int tid, int nthr) {
  constexpr int U = D / 8;                       // 16-byte units per row
  for (int i = tid; i < nrows * U; i += nthr) {
    const int r = i / U, c = (i - r * U) * 8;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < nvalid) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long long)(row0 + r) * ld + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + (long long)(row0 + r) * ld + c + 4));
      o = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(dst + r * P + c) = o;
  }
}

template <int D>
__global__ void __launch_bounds__(RM_WARPS * 32)
relpos_attention_mma_kernel(const float* __restrict__ qkv, long long ld, const float* __restrict__ relk,
                            const float* __restrict__ relv, int window, int T, int H,
                            const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int P = D + 8;                       // padded pitch (halfs): conflict-free ldmatrix rows
  constexpr int KS = D / 16;                     // k-steps of Q K^T
  constexpr int NT_S = RM_KT / 8;                // n8 tiles of a score tile
  constexpr int NT_O = D / 8;                    // n8 tiles of the output
  extern __shared__ __align__(16) unsigned char smraw[];
  __half* qs = reinterpret_cast<__half*>(smraw);            // [64][P]   (also: Ek, padded to 16 rows, during the prologue)
  __half* ks = qs + RM_QT * P;                              // [64][P]
  __half* vs = ks + RM_KT * P;                              // [64][P]
  float* rb = reinterpret_cast<float*>(vs + RM_KT * P);     // [64][16]  q . Ek[r]
  float* band = rb + RM_QT * 16;                            // [64][12]  scaled scores inside the +-w band
  float* evs = band + RM_QT * 12;                           // [9][D]    rel-value embeddings (fp32)
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * RM_QT;
  const int len = lens ? min(lens[b], T) : T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int nrel = 2 * window + 1;
  const int HD = H * D;
  const float* base = qkv + (long long)b * T * ld + h * D;

  if (q0 >= len) {                               // the whole query block is padding
    for (int i = tid; i < RM_QT * D; i += blockDim.x) {
      const int qi = q0 + i / D;
      if (qi < T) stany(out, ((long long)b * T + qi) * out_ld + h * D + (i % D), 0.f, odt);
    }
    return;
  }

  // ---- prologue: Q tile, Ek (as 16 "keys" in the K buffer), Ev, band = -inf ----
  stage_f16<D, P>(qs, base, ld, q0, RM_QT, len, tid, blockDim.x);
  stage_f16<D, P>(ks, relk, D, 0, 16, nrel, tid, blockDim.x);
  for (int i = tid; i < RM_NREL * D; i += blockDim.x) evs[i] = i < nrel * D ? relv[i] : 0.f;
  for (int i = tid; i < RM_QT * 12; i += blockDim.x) band[i] = -INFINITY;
  __syncthreads();

  // this warp's Q fragments (16 queries x D), kept in registers for the whole kernel
  uint32_t qa[KS][4];
  {
    const uint32_t a0 = smem_u32(qs + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * P + (lane >> 4) * 8);
#pragma unroll
    for (int k = 0; k < KS; ++k) ldsm_x4(a0 + k * 32, qa[k][0], qa[k][1], qa[k][2], qa[k][3]);
  }
  // rel-key side GEMM: rb[q][r] = q . Ek[r]
  {
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t b0 = smem_u32(ks + ((lane & 7) + (lane >> 4) * 8) * P + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      uint32_t r0, r1, r2, r3;
      ldsm_x4(b0 + k * 32, r0, r1, r2, r3);
      mma16816(c0, qa[k], r0, r1);
      mma16816(c1, qa[k], r2, r3);
    }
    float* r = rb + (warp * 16 + g) * 16 + 2 * t4;
    r[0] = c0[0]; r[1] = c0[1]; r[8] = c1[0]; r[9] = c1[1];
    r[128] = c0[2]; r[129] = c0[3]; r[136] = c1[2]; r[137] = c1[3];
  }
  __syncthreads();                               // Ek fragments consumed: the K buffer may be overwritten

  float oacc[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float scale = rsqrtf((float)D);
  const int qw0 = q0 + warp * 16;                // first query of this warp
  const int qrow[2] = {qw0 + g, qw0 + g + 8};

  const int ntiles = (len + RM_KT - 1) / RM_KT;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int j0 = kt * RM_KT;
    stage_f16<D, P>(ks, base + HD, ld, j0, RM_KT, len, tid, blockDim.x);
    stage_f16<D, P>(vs, base + 2 * HD, ld, j0, RM_KT, len, tid, blockDim.x);
    __syncthreads();

    // S = Q K^T (16 x 64 per warp)
    float s[NT_S][4];
#pragma unroll
    for (int n = 0; n < NT_S; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    const uint32_t kb = smem_u32(ks + ((lane & 7) + (lane >> 4) * 8) * P + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
#pragma unroll
      for (int n2 = 0; n2 < NT_S / 2; ++n2) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(kb + (n2 * 16 * P) * 2 + k * 32, r0, r1, r2, r3);
        mma16816(s[2 * n2], qa[k], r0, r1);
        mma16816(s[2 * n2 + 1], qa[k], r2, r3);
      }
    }
    // rel-key bias inside the band, scale, mask; park the band's scores for the rel-value term
    const bool near_diag = j0 <= qw0 + 15 + window && j0 + RM_KT - 1 >= qw0 - window;     // warp-uniform
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + n * 8 + 2 * t4 + (e & 1);
        const int rr = e >> 1;
        float v = s[n][e];
        if (near_diag) {
          const int rel = j - qrow[rr];
          if (rel >= -window && rel <= window) {
            const int ql = qrow[rr] - q0;
            v += rb[ql * 16 + rel + window];
            v *= scale;
            if (j < len) band[ql * 12 + rel + window] = v;
            s[n][e] = j < len ? v : -INFINITY;
            continue;
          }
        }
        s[n][e] = j < len ? v * scale : -INFINITY;
      }
    }
    // online softmax (rows g and g + 8 of the warp's 16 queries; a row lives in the 4 lanes of a quad)
    float alpha[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NT_S; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * rr], s[n][2 * rr + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[rr], mx);          // finite: tile 0 always holds key 0 < len
      alpha[rr] = __expf(m_run[rr] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < NT_S; ++n) {
        s[n][2 * rr] = __expf(s[n][2 * rr] - m_new);
        s[n][2 * rr + 1] = __expf(s[n][2 * rr + 1] - m_new);
        sum += s[n][2 * rr] + s[n][2 * rr + 1];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      l_run[rr] = l_run[rr] * alpha[rr] + sum;
      m_run[rr] = m_new;
    }
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      oacc[n][0] *= alpha[0]; oacc[n][1] *= alpha[0];
      oacc[n][2] *= alpha[1]; oacc[n][3] *= alpha[1];
    }
    // O += P V  (P: the score fragments re-packed as fp16 A operands; V through transposing ldmatrix)
    const uint32_t vb = smem_u32(vs + ((lane & 7) + ((lane >> 3) & 1) * 8) * P + (lane >> 4) * 8);
#pragma unroll
    for (int kk = 0; kk < RM_KT / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < NT_O / 2; ++n2) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(vb + (kk * 16 * P) * 2 + n2 * 32, r0, r1, r2, r3);
        mma16816(oacc[2 * n2], pa, r0, r1);
        mma16816(oacc[2 * n2 + 1], pa, r2, r3);
      }
    }
    __syncthreads();                             // everyone is done with this K / V tile (and its band writes landed)
  }

  // ---- epilogue: normalise, add the rel-value term, store ----
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int qi = qrow[rr];
    if (qi >= T) continue;
    const bool valid = qi < len;
    const float inv = valid ? 1.f / l_run[rr] : 0.f;
    float pb[RM_NREL];
#pragma unroll
    for (int r = 0; r < RM_NREL; ++r) {
      const float sv = band[(qi - q0) * 12 + r];
      pb[r] = (valid && r < nrel) ? __expf(sv - m_run[rr]) * inv : 0.f;      // exp(-inf) = 0 outside the sequence
    }
    const long long orow = ((long long)b * T + qi) * out_ld + h * D;
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      const int c = n * 8 + 2 * t4;
      float o0 = oacc[n][2 * rr] * inv, o1 = oacc[n][2 * rr + 1] * inv;
#pragma unroll
      for (int r = 0; r < RM_NREL; ++r) {
        const float2 ev = *reinterpret_cast<const float2*>(evs + r * D + c);
        o0 += pb[r] * ev.x; o1 += pb[r] * ev.y;
      }
      if (odt == AS_F32) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + orow + c) = make_float2(o0, o1);
      } else {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(out) + orow + c) = pack16(o0, o1, odt);
      }
    }
  }
}

size_t relpos_mma_smem(int D) {
  const size_t P = D + 8;
  return (size_t)(RM_QT + 2 * RM_KT) * P * 2 + sizeof(float) * (RM_QT * 16 + RM_QT * 12 + RM_NREL * D);
}

int relpos_attention_mma_launch(const float* qkv, long long ld, const float* relk, const float* relv, int window, int B,
                                int T, int H, const int* lens, void* out, int odt, long long out_ld, cudaStream_t st) {
  const size_t smem = relpos_mma_smem(128);
  ASB_SMEM_OPT_IN(100 * 1024, relpos_attention_mma_kernel<128>);
  dim3 grid((T + RM_QT - 1) / RM_QT, H, B);
  ASB_CUDA(launch_k(relpos_attention_mma_kernel<128>, grid, RM_WARPS * 32, smem, st, qkv, ld, relk, relv, window, T, H, lens, out,
                    odt, out_ld));
  return AS_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Conformer (Transformer-XL) attention on tensor cores (Utils/EMA/conformer/conformer/attention.py:72-113):
//   score[a][j] = ((q_a + u) . k_j + shift((q + v) . pos^T)[a][j]) / sqrt(d_model),   out_a = softmax_j(score) . v
// The reference's view-based relative shift reads the position scores M = (Q + v) P^T at
//   j <= a: M[a][L - a + j - 1],   j == a + 1: 0,   j >= a + 2: M[a + 1][j - a - 2]        (L = the item's own length)
// i.e. a row-dependent column shift that crosses fragment boundaries, so M goes through shared memory: phase 1 computes
// the 64 rows a0 .. a0 + 63 of M for all L columns with mma.sync (one m16 tile per warp), phase 2 runs the flash-style
// pass over the 48 queries a0 .. a0 + 47 (warps 0-2; row a0 + 48 of M is the "a + 1" of the last query).
// ---------------------------------------------------------------------------------------------------------------
constexpr int CM_QT = 48;        // queries per CTA
constexpr int CM_MR = 64;        // rows of M per CTA
constexpr int CM_KT = 64;        // keys / positions per tile
constexpr int CM_MAXT = 448;     // longest sequence whose M rows fit in shared memory

template <int D>
__global__ void __launch_bounds__(128)
conformer_attention_mma_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                               long long ld, const float* __restrict__ pos, const float* __restrict__ ub,
                               const float* __restrict__ vb, int T, int H, const int* __restrict__ lens, void* out,
                               int odt, long long out_ld) {
  pdl_wait();
  constexpr int P = D + 8;
  constexpr int KS = D / 16;
  constexpr int NT_S = CM_KT / 8;
  constexpr int NT_O = D / 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  __half* qu = reinterpret_cast<__half*>(smraw);           // [48][P]  q + u
  __half* qv = qu + CM_QT * P;                             // [64][P]  q + v
  __half* ks = qv + CM_MR * P;                             // [64][P]  key / position tile
  __half* vs = ks + CM_KT * P;                             // [64][P]  value tile
  float* Ms = reinterpret_cast<float*>(vs + CM_KT * P);    // [64][Tpad] position scores
  const int b = blockIdx.z, h = blockIdx.y, a0 = blockIdx.x * CM_QT;
  const int L = lens ? min(lens[b], T) : T;
  const int Tpad = ((T + 63) & ~63) + 4;                   // +4 floats: rows land in different banks
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int HD = H * D;
  if (a0 >= L) {
    for (int i = tid; i < CM_QT * D; i += blockDim.x) {
      const int r = a0 + i / D;
      if (r < T) stany(out, ((long long)b * T + r) * out_ld + h * D + (i % D), 0.f, odt);
    }
    return;
  }
  const float* qb = q + (long long)b * T * ld + h * D;
  const float* kb = k + (long long)b * T * ld + h * D;
  const float* vbase = v + (long long)b * T * ld + h * D;
  // q + u / q + v in fp16 (rows beyond the item's length carry the bias only, as in the reference's padded rows)
  for (int i = tid; i < CM_MR * (D / 2); i += blockDim.x) {
    const int r = i / (D / 2), d = (i - r * (D / 2)) * 2, a = a0 + r;
    float2 x = make_float2(0.f, 0.f);
    if (a < L) x = *reinterpret_cast<const float2*>(qb + (long long)a * ld + d);
    const float2 bu = *reinterpret_cast<const float2*>(ub + h * D + d), bv = *reinterpret_cast<const float2*>(vb + h * D + d);
    if (r < CM_QT) *reinterpret_cast<uint32_t*>(qu + r * P + d) = pack_h2(x.x + bu.x, x.y + bu.y);
    *reinterpret_cast<uint32_t*>(qv + r * P + d) = pack_h2(x.x + bv.x, x.y + bv.y);
  }
  __syncthreads();
  const int frag_row = (lane & 7) + ((lane >> 3) & 1) * 8, frag_col = (lane >> 4) * 8;       // A-operand ldmatrix lane map
  const uint32_t b_off = (uint32_t)((((lane & 7) + (lane >> 4) * 8) * P + ((lane >> 3) & 1) * 8) * 2);   // B (K-major rows)

  // ---- phase 1: M[a0 + 16 * warp + (0..15)][0..L) = (q + v) . pos^T ----
  {
    uint32_t av[KS][4];
    const uint32_t a_addr = smem_u32(qv + (warp * 16 + frag_row) * P + frag_col);
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) ldsm_x4(a_addr + kk * 32, av[kk][0], av[kk][1], av[kk][2], av[kk][3]);
    for (int j0 = 0; j0 < L; j0 += CM_KT) {
      __syncthreads();
      stage_f16<D, P>(ks, pos + h * D, HD, j0, CM_KT, L, tid, blockDim.x);
      __syncthreads();
      float m[NT_S][4];
#pragma unroll
      for (int n = 0; n < NT_S; ++n) m[n][0] = m[n][1] = m[n][2] = m[n][3] = 0.f;
      const uint32_t kaddr = smem_u32(ks) + b_off;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
        for (int n2 = 0; n2 < NT_S / 2; ++n2) {
          uint32_t r0, r1, r2, r3;
          ldsm_x4(kaddr + (n2 * 16 * P) * 2 + kk * 32, r0, r1, r2, r3);
          mma16816(m[2 * n2], av[kk], r0, r1);
          mma16816(m[2 * n2 + 1], av[kk], r2, r3);
        }
      }
      float* mr = Ms + (warp * 16 + g) * Tpad + j0 + 2 * t4;
#pragma unroll
      for (int n = 0; n < NT_S; ++n) {
        *reinterpret_cast<float2*>(mr + n * 8) = make_float2(m[n][0], m[n][1]);
        *reinterpret_cast<float2*>(mr + 8 * Tpad + n * 8) = make_float2(m[n][2], m[n][3]);
      }
    }
  }
  __syncthreads();

  // ---- phase 2: flash pass over the keys for the 48 queries (warps 0-2) ----
  uint32_t au[KS][4];
  if (warp < 3) {
    const uint32_t a_addr = smem_u32(qu + (warp * 16 + frag_row) * P + frag_col);
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) ldsm_x4(a_addr + kk * 32, au[kk][0], au[kk][1], au[kk][2], au[kk][3]);
  }
  float oacc[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float inv_sqrt = rsqrtf((float)HD);
  const int arow[2] = {a0 + warp * 16 + g, a0 + warp * 16 + g + 8};
  for (int j0 = 0; j0 < L; j0 += CM_KT) {
    __syncthreads();
    stage_f16<D, P>(ks, kb, ld, j0, CM_KT, L, tid, blockDim.x);
    stage_f16<D, P>(vs, vbase, ld, j0, CM_KT, L, tid, blockDim.x);
    __syncthreads();
    if (warp >= 3) continue;                     // warp 3 only helps staging in this phase
    float s[NT_S][4];
#pragma unroll
    for (int n = 0; n < NT_S; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    const uint32_t kaddr = smem_u32(ks) + b_off;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < NT_S / 2; ++n2) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(kaddr + (n2 * 16 * P) * 2 + kk * 32, r0, r1, r2, r3);
        mma16816(s[2 * n2], au[kk], r0, r1);
        mma16816(s[2 * n2 + 1], au[kk], r2, r3);
      }
    }
    // + shifted position scores, scale, mask keys >= L
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + n * 8 + 2 * t4 + (e & 1);
        const int a = arow[e >> 1];
        float x = s[n][e];
        if (j < L) {
          const int rl = a - a0;                 // row of M relative to the CTA
          if (a < L) {                           // (padding queries: their rows are discarded)
            if (j <= a) x += Ms[rl * Tpad + (L - a + j - 1)];
            else if (j >= a + 2) x += Ms[(rl + 1) * Tpad + (j - a - 2)];
          }
          x *= inv_sqrt;
        } else {
          x = -INFINITY;
        }
        s[n][e] = x;
      }
    }
    float alpha[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NT_S; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * rr], s[n][2 * rr + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[rr], mx);
      alpha[rr] = __expf(m_run[rr] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < NT_S; ++n) {
        s[n][2 * rr] = __expf(s[n][2 * rr] - m_new);
        s[n][2 * rr + 1] = __expf(s[n][2 * rr + 1] - m_new);
        sum += s[n][2 * rr] + s[n][2 * rr + 1];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      l_run[rr] = l_run[rr] * alpha[rr] + sum;
      m_run[rr] = m_new;
    }
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      oacc[n][0] *= alpha[0]; oacc[n][1] *= alpha[0];
      oacc[n][2] *= alpha[1]; oacc[n][3] *= alpha[1];
    }
    const uint32_t vaddr = smem_u32(vs + frag_row * P + frag_col);
#pragma unroll
    for (int kk = 0; kk < CM_KT / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < NT_O / 2; ++n2) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(vaddr + (kk * 16 * P) * 2 + n2 * 32, r0, r1, r2, r3);
        mma16816(oacc[2 * n2], pa, r0, r1);
        mma16816(oacc[2 * n2 + 1], pa, r2, r3);
      }
    }
  }
  if (warp >= 3) return;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int a = arow[rr];
    if (a >= T) continue;
    const float inv = a < L ? 1.f / l_run[rr] : 0.f;
    const long long orow = ((long long)b * T + a) * out_ld + h * D;
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      const int c = n * 8 + 2 * t4;
      const float o0 = oacc[n][2 * rr] * inv, o1 = oacc[n][2 * rr + 1] * inv;
      if (odt == AS_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + orow + c) = make_float2(o0, o1);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(out) + orow + c) = pack16(o0, o1, odt);
    }
  }
}

bool conformer_mma_eligible(int T) { return T <= CM_MAXT; }

int conformer_attention_mma_launch(const float* q, const float* k, const float* v, long long ld, const float* pos,
                                   const float* ub, const float* vb, int B, int T, int H, const int* lens, void* out,
                                   int odt, long long out_ld, cudaStream_t st) {
  constexpr int D = 64, P = D + 8;
  const int Tpad = ((T + 63) & ~63) + 4;
  const size_t smem = (size_t)(CM_QT + CM_MR + 2 * CM_KT) * P * 2 + sizeof(float) * (size_t)CM_MR * Tpad;
  ASB_SMEM_OPT_IN(200 * 1024, conformer_attention_mma_kernel<64>);
  dim3 grid((T + CM_QT - 1) / CM_QT, H, B);
  ASB_CUDA(launch_k(conformer_attention_mma_kernel<64>, grid, 128, smem, st, q, k, v, ld, pos, ub, vb, T, H, lens, out, odt, out_ld));
  return AS_OK;
}

}  // namespace asb
