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

}  // namespace asb

using namespace asb;

extern "C" int as_bilstm(const float* xproj, int64_t xproj_ld, const float* whh, int32_t B, int32_t T,
                         int32_t H, const int32_t* lens, void* out, int32_t out_dtype,
                         int64_t out_ld, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(xproj && whh && out, AS_ERR_SHAPE, "as_bilstm: null pointer");
  ASB_REQUIRE(H == 128 || H == 256, AS_ERR_SHAPE, "as_bilstm: hidden size %d unsupported (128 or 256)", H);
  const int G = 4 * H;
  const size_t smem = sizeof(float) * ((size_t)2 * LSTM_BG * H + (size_t)LSTM_BG * G);
  dim3 grid((B + LSTM_BG - 1) / LSTM_BG, 2);
  bilstm_kernel<LSTM_BG><<<grid, G, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      xproj, xproj_ld, whh, B, T, H, lens, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
