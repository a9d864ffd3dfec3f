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

template <int BG>
__global__ void __launch_bounds__(1024, 1) bilstm_kernel(const float* __restrict__ xproj, long long xp_ld,
                              const float* __restrict__ whh_t /* [2][H][4H] */, int B, int T, int H,
                              const int* __restrict__ lens, void* out, int odt, long long out_ld) {
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

  auto xp_at = [&](int n, int len, int s, int q, int j) -> float {
    if (s >= len) return 0.f;
    const int t = dir == 0 ? s : len - 1 - s;
    return __ldg(xproj + ((long long)(b0 + n) * T + t) * xp_ld + (long long)dir * G + q * H + j);
  };
  float xn[4][4];  // prefetched input projections for the next step: [gate][element]
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xn[q][0] = xp_at(n0, len0, 0, q, j0); xn[q][1] = xp_at(n0 + 1, len1, 0, q, j0);
    xn[q][2] = xp_at(n0, len0, 0, q, j1); xn[q][3] = xp_at(n0 + 1, len1, 0, q, j1);
  }

  for (int s = 0; s < maxlen; ++s) {
    const __half* hp = &hs[s & 1][0][0];
    __half* hn = &hs[(s + 1) & 1][0][0];
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = xn[q][e];
    // prefetch step s+1 while the tensor cores work
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      xn[q][0] = xp_at(n0, len0, s + 1, q, j0); xn[q][1] = xp_at(n0 + 1, len1, s + 1, q, j0);
      xn[q][2] = xp_at(n0, len0, s + 1, q, j1); xn[q][3] = xp_at(n0 + 1, len1, s + 1, q, j1);
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // B fragment: B[k][n] = h[n][k]; b0 -> k = 16ks + 2tig (+1), b1 -> +8, n = gid
      const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(hp + gid * L128_HP + 16 * ks + 2 * tig);
      const uint32_t bb1 = *reinterpret_cast<const uint32_t*>(hp + gid * L128_HP + 16 * ks + 2 * tig + 8);
#pragma unroll
      for (int q = 0; q < 4; ++q) mma16816(acc[q], afrag[q][ks], bb0, bb1);
    }
    // cell update in registers (fp32): element e -> (unit, batch) = (j0,n0) (j0,n0+1) (j1,n0) (j1,n0+1)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + (e & 1), j = (e < 2) ? j0 : j1;
      const int len = (e & 1) ? len1 : len0;
      if (s < len) {
        const float ig = sigm(acc[0][e]), fg = sigm(acc[1][e]), gg = tanhf(acc[2][e]), og = sigm(acc[3][e]);
        c[e] = fg * c[e] + ig * gg;
        const float hval = og * tanhf(c[e]);
        hn[n * L128_HP + j] = __float2half_rn(hval);
        const int t = dir == 0 ? s : len - 1 - s;
        stany(out, ((long long)(b0 + n) * T + t) * out_ld + dir * H + j, hval, odt);
      }
    }
    __syncthreads();
  }
}

}  // namespace asb

using namespace asb;

extern "C" int as_bilstm(const float* xproj, int64_t xproj_ld, const float* whh, int32_t B, int32_t T,
                         int32_t H, const int32_t* lens, void* out, int32_t out_dtype,
                         int64_t out_ld, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(xproj && whh && out, AS_ERR_SHAPE, "as_bilstm: null pointer");
  ASB_REQUIRE(H == 128 || H == 256, AS_ERR_SHAPE, "as_bilstm: hidden size %d unsupported (128 or 256)", H);
  if (H == 128) {
    dim3 grid128((B + L128_NB - 1) / L128_NB, 2);
    bilstm128_mma_kernel<<<grid128, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xproj, xproj_ld, whh, B, T, lens, out, out_dtype, out_ld);
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  const int G = 4 * H;
  const size_t smem = sizeof(float) * ((size_t)2 * LSTM_BG * H + (size_t)LSTM_BG * G);
  dim3 grid((B + LSTM_BG - 1) / LSTM_BG, 2);
  bilstm_kernel<LSTM_BG><<<grid, G, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      xproj, xproj_ld, whh, B, T, H, lens, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
