// Single-output-channel 1-D convolution (the vocoder's conv_post, Vocoder/vocoder.py:113-114:
// Conv1d(32, 1, 7) + tanh on 3.84 M positions per batch of 16 x 10 s).  As an implicit GEMM it wastes a
// 128 x 16 tensor-core tile on one useful column and took 209 us; it is a pure HBM stream (read 64 B,
// write 4 B per position): one CTA stages 256 + k - 1 rows in shared memory, one thread per output.
#include "common.cuh"

namespace asb {

constexpr int C1_ROWS = 256;

template <bool BF16>
__global__ void __launch_bounds__(C1_ROWS)
conv_cout1_kernel(const uint16_t* __restrict__ x, long long x_ld, int T, int Cin, const uint16_t* __restrict__ w,
                  long long w_tap_stride, int ntaps, int dt0, int dil, const float* __restrict__ bias, const int* __restrict__ lens,
                  float out_scale, int act, float slope, void* y_raw, int y_raw_dtype, long long y_raw_ld,
                  void* y_act, int y_act_dtype, long long y_act_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(16) uint16_t tile[];        // [(C1_ROWS + halo)][Cin] 16-bit, then weights fp32 [ntaps][Cin]
  const int b = blockIdx.y, t0 = blockIdx.x * C1_ROWS;
  const int halo = (ntaps - 1) * dil;
  const int nrows = C1_ROWS + halo;
  const int upr = Cin / 8;                                // 16-byte units per row
  float* ws = reinterpret_cast<float*>(tile + (size_t)nrows * Cin);
  for (int i = threadIdx.x; i < ntaps * Cin; i += blockDim.x) {
    const uint16_t u = w[(long long)(i / Cin) * w_tap_stride + (i % Cin)];
    ws[i] = BF16 ? __uint_as_float((uint32_t)u << 16) : __half2float(__ushort_as_half(u));
  }
  const uint16_t* xb = x + (long long)b * T * x_ld;
  for (int i = threadIdx.x; i < nrows * upr; i += blockDim.x) {
    const int r = i / upr, u = i - r * upr;
    const int t = t0 + dt0 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const uint4*>(xb + (long long)t * x_ld) + u);
    // rotate the 16-byte units of a row by the row index: threads of a warp (consecutive rows, same unit)
    // then hit different banks
    reinterpret_cast<uint4*>(tile)[r * upr + ((u + r) % upr)] = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float acc = bias != nullptr ? bias[0] : 0.f;
  for (int j = 0; j < ntaps; ++j) {
    const int r = threadIdx.x + j * dil;
    const float* wj = ws + j * Cin;
    for (int u = 0; u < upr; ++u) {
      const uint4 v = reinterpret_cast<const uint4*>(tile)[r * upr + ((u + r) % upr)];
      const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a0, a1;
        if (BF16) { a0 = __uint_as_float(q[e] << 16); a1 = __uint_as_float(q[e] & 0xFFFF0000u); }
        else { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[e])); a0 = f.x; a1 = f.y; }
        acc += a0 * wj[u * 8 + 2 * e] + a1 * wj[u * 8 + 2 * e + 1];
      }
    }
  }
  acc *= out_scale;
  if (lens != nullptr && t >= lens[b]) acc = 0.f;
  const long long row = (long long)b * T + t;
  if (y_raw) stany(y_raw, row * y_raw_ld, acc, y_raw_dtype);
  if (y_act) stany(y_act, row * y_act_ld, apply_act(acc, act, slope), y_act_dtype);
}

// eligible: 1-D, one output channel, 16-bit input with Cin % 8 == 0 and Cin <= 64, equally spaced taps, no residuals
bool conv_cout1_eligible(const as_conv_params* p) {
  if (p->Cout != 1 || p->F != 1 || p->Fo != 1 || p->To != p->T || p->res1 || p->res2 || p->stats) return false;
  if (p->Cin % 8 != 0 || p->Cin > 64 || p->ntaps > 16 || (p->x_ld % 8) != 0) return false;
  const int dil = p->ntaps > 1 ? p->tap_dt[1] - p->tap_dt[0] : 1;
  if (dil < 1) return false;
  for (int j = 0; j < p->ntaps; ++j)
    if (p->tap_df[j] != 0 || p->tap_dt[j] != p->tap_dt[0] + j * dil) return false;
  return (p->ntaps - 1) * dil <= 64;
}

int conv_cout1_launch(const as_conv_params* p, cudaStream_t st) {
  const int dil = p->ntaps > 1 ? p->tap_dt[1] - p->tap_dt[0] : 1;
  const int nrows = C1_ROWS + (p->ntaps - 1) * dil;
  const size_t smem = (size_t)nrows * p->Cin * 2 + (size_t)p->ntaps * p->Cin * 4;
  dim3 grid((unsigned)((p->T + C1_ROWS - 1) / C1_ROWS), (unsigned)p->B);
  const long long w_tap_stride = (long long)p->CoutP * p->CinP;    // row 0 (the only output channel) of every tap
  if (p->x_dtype == AS_BF16)
    ASB_CUDA(launch_k(conv_cout1_kernel<true>, grid, C1_ROWS, smem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
        reinterpret_cast<const uint16_t*>(p->w), w_tap_stride, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
        p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
  else
    ASB_CUDA(launch_k(conv_cout1_kernel<false>, grid, C1_ROWS, smem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
        reinterpret_cast<const uint16_t*>(p->w), w_tap_stride, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
        p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

}  // namespace asb
