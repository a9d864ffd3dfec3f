// Bidirectional LSTM recurrence (PyTorch gate order i,f,g,o) with packed-sequence semantics.
// Replaces nn.LSTM in DurationPredictor (models.py:555-564), ArtsPredictor (:606-618) and JDCNet
// (Utils/JDC/model.py:128).  The input projections W_ih x + b_ih + b_hh for both directions are
// ONE implicit-GEMM launch done beforehand; this kernel is only the serial part.
//
// One persistent CTA per (direction, group of BG batch items): 4H threads, thread r owns gate row
// r.  Each step: gate[r][b] = xproj[b][t][r] + sum_k Whh_T[k][r] * h[b][k]  (Whh_T is k-major so
// the 4H threads read 4H consecutive floats: coalesced, L1/L2 resident across steps), then the
// first H*BG threads apply the cell update in fp32 and publish h through shared memory.
// Everything (weights, state, accumulation) is fp32.
#include "common.cuh"

namespace asb {

constexpr int LSTM_BG = 4;  // batch items per CTA

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }
// MUFU-based variants for the per-step critical path of the tensor-core kernels (abs error ~1e-6)
__device__ __forceinline__ float fsigm(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
__device__ __forceinline__ float ftanh(float v) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * v)); }
// bare MUFU forms (no range fix-ups: ex2 -> +inf gives rcp -> 0, the correct limit; abs error ~3e-7)
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gsigm(float v) { return mufu_rcp(1.f + mufu_ex2(-1.4426950408889634f * v)); }
__device__ __forceinline__ float gtanh(float v) { return fmaf(-2.f, mufu_rcp(1.f + mufu_ex2(2.8853900817779268f * v)), 1.f); }

template <int BG>
__global__ void __launch_bounds__(1024, 1) bilstm_kernel(const float* __restrict__ xproj, long long xp_ld,
                              const float* __restrict__ whh_t /* [2][H][4H] */, int B, int T, int H,
                              const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ float sm[];
  const int G = 4 * H;
  float* hbuf = sm;                 // [2][BG][H] ping-pong
  float* gates = hbuf + 2 * BG * H; // [BG][4H]
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BG;
  const int r = threadIdx.x;        // gate row
  const float* W = whh_t + (long long)dir * H * G;

  int len[BG];
  int maxlen = 0;
#pragma unroll
  for (int i = 0; i < BG; ++i) {
    const int b = b0 + i;
    len[i] = (b < B) ? (lens ? min(lens[b], T) : T) : 0;
    maxlen = max(maxlen, len[i]);
  }
  for (int i = r; i < 2 * BG * H; i += blockDim.x) hbuf[i] = 0.f;
  // cell state: thread (i = r / H, j = r % H) for r < BG*H  (requires BG <= 4)
  float c_state = 0.f;
  const int ui = r / H, uj = r % H;
  __syncthreads();

  // zero the padded tail of the output (pad_packed_sequence semantics)
  for (int i = 0; i < BG; ++i) {
    const int b = b0 + i;
    if (b >= B) continue;
    for (long long e = (long long)len[i] * H + r; e < (long long)T * H; e += blockDim.x) {
      const int t = e / H, j = e % H;
      stany(out, ((long long)b * T + t) * out_ld + dir * H + j, 0.f, odt);
    }
  }

  for (int s = 0; s < maxlen; ++s) {
    const float* hprev = hbuf + (s & 1) * BG * H;
    float* hnext = hbuf + ((s + 1) & 1) * BG * H;
    float acc[BG];
#pragma unroll
    for (int i = 0; i < BG; ++i) {
      // time index of this step for item i: forward t = s, reverse t = len-1-s
      const int t = dir == 0 ? s : len[i] - 1 - s;
      const int b = b0 + i;
      acc[i] = (s < len[i]) ? xproj[((long long)b * T + t) * xp_ld + (long long)dir * G + r] : 0.f;
    }
#pragma unroll 4
    for (int k = 0; k < H; k += 4) {
      const float w0 = W[(long long)(k + 0) * G + r];
      const float w1 = W[(long long)(k + 1) * G + r];
      const float w2 = W[(long long)(k + 2) * G + r];
      const float w3 = W[(long long)(k + 3) * G + r];
#pragma unroll
      for (int i = 0; i < BG; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(hprev + i * H + k);
        acc[i] += w0 * hv.x + w1 * hv.y + w2 * hv.z + w3 * hv.w;
      }
    }
#pragma unroll
    for (int i = 0; i < BG; ++i) gates[i * G + r] = acc[i];
    __syncthreads();
    if (r < BG * H) {
      const int b = b0 + ui;
      if (s < len[ui]) {
        const float* g = gates + ui * G;
        const float ig = sigm(g[uj]), fg = sigm(g[H + uj]), gg = tanhf(g[2 * H + uj]), og = sigm(g[3 * H + uj]);
        c_state = fg * c_state + ig * gg;
        const float hval = og * tanhf(c_state);
        hnext[ui * H + uj] = hval;
        const int t = dir == 0 ? s : len[ui] - 1 - s;
        stany(out, ((long long)b * T + t) * out_ld + dir * H + uj, hval, odt);
      } else {
        hnext[ui * H + uj] = hprev[ui * H + uj];
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// H = 128 fast path (the three ArtsPredictor LSTMs run Tm = 800 .. 4800 sequential steps).
// Per step the recurrent product gates[512 x 8] = W_hh[512 x 128] . h[128 x 8] runs on mma.sync
// m16n8k16 (fp16 operands, fp32 accumulate) with the W_hh fragments RESIDENT IN REGISTERS for the
// whole sequence (128 regs/thread), so a step touches no weight memory at all.  Warp w owns hidden
// units 16w..16w+15 for all four gates, which puts i,f,g,o of one (unit, batch item) in the same
// thread: the cell update is register-only (fp32 c and h), h is re-published as fp16 through a
// double-buffered shared tile (one __syncthreads per step).  8 batch items per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int L128_H = 128;
constexpr int L128_NB = 8;     // batch items per CTA (mma N)
constexpr int L128_HP = 136;   // padded pitch (halves) of the h tile: conflict-free fragment loads

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 1)
bilstm128_mma_kernel(const float* __restrict__ xproj, long long xp_ld,
                     const float* __restrict__ whh_t /* [2][H][4H] */, int B, int T,
                     const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int H = L128_H, G = 4 * H;
  __shared__ __align__(16) __half hs[2][L128_NB][L128_HP];
  __shared__ int s_len[L128_NB];
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * L128_NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const float* Wt = whh_t + (long long)dir * H * G;      // Wt[k][row]

  if (threadIdx.x < L128_NB) {
    const int b = b0 + threadIdx.x;
    s_len[threadIdx.x] = (b < B) ? (lens ? min(lens[b], T) : T) : 0;
  }
  for (int i = threadIdx.x; i < 2 * L128_NB * L128_HP; i += blockDim.x) (&hs[0][0][0])[i] = __float2half(0.f);
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int i = 0; i < L128_NB; ++i) maxlen = max(maxlen, s_len[i]);

  // A fragments: tile q (gate) rows = q*H + 16*warp + m; k-step ks covers k = 16*ks .. 16*ks+15
  uint32_t afrag[4][8][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r_lo = q * H + 16 * warp + gid, r_hi = r_lo + 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int k0 = 16 * ks + 2 * tig;
      auto pk = [&](int r, int k) {
        __half2 v = __floats2half2_rn(Wt[(long long)k * G + r], Wt[(long long)(k + 1) * G + r]);
        return *reinterpret_cast<uint32_t*>(&v);
      };
      afrag[q][ks][0] = pk(r_lo, k0);
      afrag[q][ks][1] = pk(r_hi, k0);
      afrag[q][ks][2] = pk(r_lo, k0 + 8);
      afrag[q][ks][3] = pk(r_hi, k0 + 8);
    }
  }

  // this thread's accumulator elements: hidden units j0 (rows gid) / j1 (rows gid+8), batch n0, n0+1
  const int j0 = 16 * warp + gid, j1 = j0 + 8;
  const int n0 = 2 * tig;
  const int len0 = s_len[n0], len1 = s_len[n0 + 1];
  float c[4] = {0.f, 0.f, 0.f, 0.f};   // (j0,n0) (j0,n1) (j1,n0) (j1,n1)

  // zero the padded tail of the output (pad_packed_sequence semantics)
  for (int i = 0; i < L128_NB; ++i) {
    const int b = b0 + i;
    if (b >= B) continue;
    for (long long e = (long long)s_len[i] * H + threadIdx.x; e < (long long)T * H; e += blockDim.x)
      stany(out, ((long long)b * T + e / H) * out_ld + dir * H + (e % H), 0.f, odt);
  }

  // Pointer-walking addressing: per thread two input rows (items n0, n0+1) and two output rows; the gate
  // (q * H) and unit (j1 = j0 + 8) offsets are compile-time immediates, the step is one pointer bump.
  // (ncu: the kernel issued ~390 instructions per warp per step, most of them 64-bit address arithmetic,
  // dtype dispatch and the range checks of __fdividef; the step is issue-bound with 2 warps per scheduler.)
  const long long step0 = dir == 0 ? xp_ld : -xp_ld, step1 = step0;
  const float* px0 = xproj + ((long long)(b0 + n0) * T + (dir == 0 ? 0 : max(len0 - 1, 0))) * xp_ld + (long long)dir * G + j0;
  const float* px1 = xproj + ((long long)(b0 + n0 + 1) * T + (dir == 0 ? 0 : max(len1 - 1, 0))) * xp_ld + (long long)dir * G + j0;
  const int es = odt == AS_F32 ? 4 : 2;
  const long long ostep = (dir == 0 ? out_ld : -out_ld) * es;
  char* po0 = reinterpret_cast<char*>(out) + (((long long)(b0 + n0) * T + (dir == 0 ? 0 : max(len0 - 1, 0))) * out_ld + dir * H + j0) * es;
  char* po1 = reinterpret_cast<char*>(out) + (((long long)(b0 + n0 + 1) * T + (dir == 0 ? 0 : max(len1 - 1, 0))) * out_ld + dir * H + j0) * es;
  // input projections are prefetched PF steps ahead (register ring; the loop is unrolled by PF so
  // the ring slots are compile-time): the L2 latency of the 16 scattered loads hides behind PF steps.
  constexpr int PF = 4;
  float xn[PF][4][4];
  auto fetch = [&](float (&dst)[4][4], int s) {
    const bool v0 = s < len0, v1 = s < len1;
    const float* p0 = px0 + (long long)s * step0;
    const float* p1 = px1 + (long long)s * step1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dst[q][0] = v0 ? __ldg(p0 + q * H) : 0.f;     dst[q][1] = v1 ? __ldg(p1 + q * H) : 0.f;
      dst[q][2] = v0 ? __ldg(p0 + q * H + 8) : 0.f; dst[q][3] = v1 ? __ldg(p1 + q * H + 8) : 0.f;
    }
  };
#pragma unroll
  for (int u = 0; u < PF; ++u) fetch(xn[u], u);
  auto put = [&](char* p, float v) {
    if (es == 4) *reinterpret_cast<float*>(p) = v;
    else *reinterpret_cast<uint16_t*>(p) = to16(v, odt);
  };

  for (int s0 = 0; s0 < maxlen; s0 += PF) {
#pragma unroll
   for (int u = 0; u < PF; ++u) {
    const int s = s0 + u;
    if (s >= maxlen) break;
    const __half* hp = &hs[s & 1][0][0];
    __half* hn = &hs[(s + 1) & 1][0][0];
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = xn[u][q][e];
    fetch(xn[u], s + PF);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // B fragment: B[k][n] = h[n][k]; b0 -> k = 16ks + 2tig (+1), b1 -> +8, n = gid
      const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(hp + gid * L128_HP + 16 * ks + 2 * tig);
      const uint32_t bb1 = *reinterpret_cast<const uint32_t*>(hp + gid * L128_HP + 16 * ks + 2 * tig + 8);
#pragma unroll
      for (int q = 0; q < 4; ++q) mma16816(acc[q], afrag[q][ks], bb0, bb1);
    }
    // cell update in registers (fp32): element e -> (unit, batch) = (j0,n0) (j0,n0+1) (j1,n0) (j1,n0+1)
    const bool live0 = s < len0, live1 = s < len1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + (e & 1), j = (e < 2) ? j0 : j1;
      if ((e & 1) ? live1 : live0) {
        const float ig = gsigm(acc[0][e]), fg = gsigm(acc[1][e]), gg = gtanh(acc[2][e]), og = gsigm(acc[3][e]);
        c[e] = fg * c[e] + ig * gg;
        const float hval = og * gtanh(c[e]);
        hn[n * L128_HP + j] = __float2half_rn(hval);
        put(((e & 1) ? po1 : po0) + (long long)s * ostep + ((e < 2) ? 0 : 8 * es), hval);
      }
    }
    __syncthreads();
   }
  }
}

// ---------------------------------------------------------------------------------------------
// H = 256 (duration predictor, JDCNet): W_hh in fp16 is 512 KB, more than one SM can hold.  A
// cluster of 4 CTAs splits the hidden units (64 each); every CTA keeps its 256 gate rows x 256 k as
// register-resident mma fragments (128 regs/thread), computes its units' gates with mma.sync
// m16n8k16 and broadcasts the new h (fp16) into all four CTAs' shared tiles through DSMEM
// (st.shared::cluster); one cluster barrier per step.  Warp w owns 8 units for all four gates:
// tile 0 = (i | f) x 8 units, tile 1 = (g | o) x 8 units, so i,f,g,o of a (unit, batch item) meet in
// one thread and the cell update is register-only (fp32).  8 batch items per cluster.
// ---------------------------------------------------------------------------------------------
// The same kernel can serve H = 128 with a cluster of 2 (64 units per CTA again); measured slower than the single-CTA
// kernel above (see as_bilstm), so it is an experiment switch only.
constexpr int L256_NB = 8;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int H, int L256_NC>
__global__ void __cluster_dims__(L256_NC, 1, 1) __launch_bounds__((H / L256_NC / 8) * 32, 1)
bilstm_cluster_kernel(const float* __restrict__ xproj, long long xp_ld,
                      const float* __restrict__ whh_t /* [2][H][4H] */, int B, int T,
                      const int* __restrict__ lens, void* out, int odt, long long out_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int UPC = H / L256_NC;         // hidden units per CTA: one warp per 8 units (64 -> 8 warps, 128 -> 16 warps)
  static_assert(UPC == 64 || UPC == 128, "8 or 16 warps per CTA");
  constexpr int G = 4 * H;
  constexpr int L256_HP = H + 8;   // padded pitch (halves) of the h tile
  constexpr int KSTEPS = H / 16;
  __shared__ __align__(16) __half hs[2][L256_NB][L256_HP];
  __shared__ int s_len[L256_NB];
  const int dir = blockIdx.y;
  const uint32_t crank = cluster_rank();
  const int b0 = (blockIdx.x / L256_NC) * L256_NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const float* Wt = whh_t + (long long)dir * H * G;      // Wt[k][row]
  const int unit = (int)crank * UPC + warp * 8 + gid;   // this thread's hidden unit

  if (threadIdx.x < L256_NB) {
    const int b = b0 + threadIdx.x;
    s_len[threadIdx.x] = (b < B) ? (lens ? min(lens[b], T) : T) : 0;
  }
  for (int i = threadIdx.x; i < 2 * L256_NB * L256_HP; i += blockDim.x) (&hs[0][0][0])[i] = __float2half(0.f);
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int i = 0; i < L256_NB; ++i) maxlen = max(maxlen, s_len[i]);

  // A fragments: tile tq rows 0..7 = gate 2tq, rows 8..15 = gate 2tq+1, both for units warp*8 + (row & 7)
  uint32_t afrag[2][KSTEPS][4];
#pragma unroll
  for (int tq = 0; tq < 2; ++tq) {
    const int r_lo = (2 * tq) * H + unit, r_hi = (2 * tq + 1) * H + unit;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int k0 = 16 * ks + 2 * tig;
      auto pk = [&](int r, int k) {
        __half2 v = __floats2half2_rn(Wt[(long long)k * G + r], Wt[(long long)(k + 1) * G + r]);
        return *reinterpret_cast<uint32_t*>(&v);
      };
      afrag[tq][ks][0] = pk(r_lo, k0);
      afrag[tq][ks][1] = pk(r_hi, k0);
      afrag[tq][ks][2] = pk(r_lo, k0 + 8);
      afrag[tq][ks][3] = pk(r_hi, k0 + 8);
    }
  }
  const int n0 = 2 * tig;
  const int len0 = s_len[n0], len1 = s_len[n0 + 1];
  float c[2] = {0.f, 0.f};   // (unit, n0), (unit, n0+1)

  // zero the padded tail of the output for this CTA's 64 units
  for (int i = 0; i < L256_NB; ++i) {
    const int b = b0 + i;
    if (b >= B) continue;
    for (long long e = (long long)s_len[i] * UPC + threadIdx.x; e < (long long)T * UPC; e += blockDim.x)
      stany(out, ((long long)b * T + e / UPC) * out_ld + dir * H + crank * UPC + (e % UPC), 0.f, odt);
  }

  auto xp_at = [&](int n, int len, int s, int gate) -> float {
    if (s >= len) return 0.f;
    const int t = dir == 0 ? s : len - 1 - s;
    return __ldg(xproj + ((long long)(b0 + n) * T + t) * xp_ld + (long long)dir * G + gate * H + unit);
  };
  constexpr int PF = 2;
  float xn[PF][4][2];   // [slot][gate][batch element]
  auto fetch = [&](float (&dst)[4][2], int s) {
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) { dst[g4][0] = xp_at(n0, len0, s, g4); dst[g4][1] = xp_at(n0 + 1, len1, s, g4); }
  };
#pragma unroll
  for (int u = 0; u < PF; ++u) fetch(xn[u], u);

  // remote addresses of this thread's h slot in each CTA of the cluster (even gid lanes publish batch
  // n0, odd gid lanes batch n0+1; each packs two adjacent units into one 32-bit DSMEM store)
  const int pub_n = (gid & 1) ? n0 + 1 : n0;
  const int pub_unit = unit & ~1;
  uint32_t raddr[L256_NC];
#pragma unroll
  for (int r = 0; r < L256_NC; ++r) raddr[r] = mapa_shared(smem_u32(&hs[0][pub_n][pub_unit]), (uint32_t)r);
  constexpr uint32_t BUF_BYTES = L256_NB * L256_HP * 2;

  auto step_sync = [&]() {      // one CTA: a block barrier; a cluster: the h slices travel through DSMEM
    if (L256_NC == 1) __syncthreads();
    else cluster_sync_all();
  };
  step_sync();   // every CTA has zeroed its tiles before anyone publishes into them

  for (int s0 = 0; s0 < maxlen; s0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int s = s0 + u;
      if (s >= maxlen) break;
      const __half* hp = &hs[s & 1][0][0];
      float acc[2][4];
      acc[0][0] = xn[u][0][0]; acc[0][1] = xn[u][0][1]; acc[0][2] = xn[u][1][0]; acc[0][3] = xn[u][1][1];
      acc[1][0] = xn[u][2][0]; acc[1][1] = xn[u][2][1]; acc[1][2] = xn[u][3][0]; acc[1][3] = xn[u][3][1];
      fetch(xn[u], s + PF);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(hp + gid * L256_HP + 16 * ks + 2 * tig);
        const uint32_t bb1 = *reinterpret_cast<const uint32_t*>(hp + gid * L256_HP + 16 * ks + 2 * tig + 8);
        mma16816(acc[0], afrag[0][ks], bb0, bb1);
        mma16816(acc[1], afrag[1][ks], bb0, bb1);
      }
      float hv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int len = e ? len1 : len0;
        hv[e] = 0.f;
        if (s < len) {
          const float ig = gsigm(acc[0][e]), fg = gsigm(acc[0][2 + e]), gg = gtanh(acc[1][e]), og = gsigm(acc[1][2 + e]);
          c[e] = fg * c[e] + ig * gg;
          hv[e] = og * gtanh(c[e]);
          const int t = dir == 0 ? s : len - 1 - s;
          stany(out, ((long long)(b0 + n0 + e) * T + t) * out_ld + dir * H + unit, hv[e], odt);
        }
      }
      // pack two adjacent units per lane pair (gid, gid^1) and publish to all CTAs of the cluster
      const float send = (gid & 1) ? hv[0] : hv[1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      __half2 pk2 = (gid & 1) ? __floats2half2_rn(recv, hv[1]) : __floats2half2_rn(hv[0], recv);
      const int plen = (gid & 1) ? len1 : len0;
      if (s < plen) {
        const uint32_t off = ((s + 1) & 1) ? BUF_BYTES : 0u;
#pragma unroll
        for (int r = 0; r < L256_NC; ++r) st_cluster_u32(raddr[r] + off, *reinterpret_cast<uint32_t*>(&pk2));
      }
      step_sync();
    }
  }
}

}  // namespace asb

using namespace asb;

extern "C" int as_bilstm(const float* xproj, int64_t xproj_ld, const float* whh, int32_t B, int32_t T,
                         int32_t H, const int32_t* lens, void* out, int32_t out_dtype,
                         int64_t out_ld, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(xproj && whh && out, AS_ERR_SHAPE, "as_bilstm: null pointer");
  ASB_REQUIRE(H == 128 || H == 256 || H == 64, AS_ERR_SHAPE, "as_bilstm: hidden size %d unsupported (64, 128 or 256)", H);
  if (H == 128) {
    // Round-2 experiments on the step time (B200, B = 16, T = 800; 8 x 4800), all SLOWER or equal to the 8-warp kernel's
    // 1.02 / 1.00 us per step, so it stays the default:
    //   * ASB_LSTM128=16warp: 16 warps in one CTA (bilstm_cluster_kernel<128, 1>, a warp per 8 hidden units, half the
    //     instructions per thread): 1.13 / 1.34 us;
    //   * ASB_LSTM128=cluster: 2-CTA cluster, h exchanged through DSMEM: 1.15 / 1.59 us (a cluster barrier per step);
    //   * two accumulator sets per gate (HMMA chain 4 deep instead of 8): 1.04 / 1.04 us.
    // ncu (profiles/r02_ncu_full_bilstm128_8warp.txt): 286 instructions per warp-step at 30 % issue utilisation, tensor
    // pipe 27 %, dominant stall "wait" -- the step is a ~2000-cycle dependency chain (operand loads -> 8 chained HMMA ->
    // MUFU chain -> h store -> barrier) that neither more warps nor a shorter HMMA chain shortened.
    static const char* mode = getenv("ASB_LSTM128");
    if (mode != nullptr && strcmp(mode, "16warp") == 0) {
      dim3 grid((B + L256_NB - 1) / L256_NB, 2);
      ASB_CUDA(launch_k(bilstm_cluster_kernel<128, 1>, grid, 512, 0, reinterpret_cast<cudaStream_t>(stream),
          xproj, xproj_ld, whh, B, T, lens, out, out_dtype, out_ld));
      return AS_OK;
    }
    if (mode != nullptr && strcmp(mode, "cluster") == 0) {
      dim3 grid(2 * ((B + L256_NB - 1) / L256_NB), 2);
      ASB_CUDA(launch_k(bilstm_cluster_kernel<128, 2>, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream),
          xproj, xproj_ld, whh, B, T, lens, out, out_dtype, out_ld));
      return AS_OK;
    }
    dim3 grid128((B + L128_NB - 1) / L128_NB, 2);
    ASB_CUDA(launch_k(bilstm128_mma_kernel, grid128, 256, 0, reinterpret_cast<cudaStream_t>(stream),
        xproj, xproj_ld, whh, B, T, lens, out, out_dtype, out_ld));
    return AS_OK;
  }
  if (H == 256) {
    dim3 grid256(4 * ((B + L256_NB - 1) / L256_NB), 2);
    ASB_CUDA(launch_k(bilstm_cluster_kernel<256, 4>, grid256, 256, 0, reinterpret_cast<cudaStream_t>(stream),
        xproj, xproj_ld, whh, B, T, lens, out, out_dtype, out_ld));
    return AS_OK;
  }
  const int G = 4 * H;
  const size_t smem = sizeof(float) * ((size_t)2 * LSTM_BG * H + (size_t)LSTM_BG * G);
  dim3 grid((B + LSTM_BG - 1) / LSTM_BG, 2);
  ASB_CUDA(launch_k(bilstm_kernel<LSTM_BG>, grid, G, smem, reinterpret_cast<cudaStream_t>(stream), 
      xproj, xproj_ld, whh, B, T, H, lens, out, out_dtype, out_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
