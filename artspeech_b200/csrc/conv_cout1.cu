// Single-output-channel 1-D convolution (the vocoder's conv_post, Vocoder/vocoder.py:113-114:
// Conv1d(32, 1, 7) + tanh on 3.84 M positions per batch of 16 x 10 s).  As an implicit GEMM it wastes a
// 128 x 16 tensor-core tile on one useful column and took 209 us; it is a pure HBM stream (read 64 B,
// write 4 B per position): one CTA stages 256 + k - 1 rows in shared memory, one thread per output.
#include "common.cuh"

namespace asb {

constexpr int C1_THREADS = 256;
constexpr int C1_OPT = 4;                               // outputs per thread (rows tid, tid + 256, ...)
constexpr int C1_ROWS = C1_THREADS * C1_OPT;            // output rows per CTA

template <bool BF16>
__global__ void __launch_bounds__(C1_THREADS)
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
    // then spread over the banks
    reinterpret_cast<uint4*>(tile)[r * upr + ((u + r) % upr)] = v;
  }
  __syncthreads();
  // each thread owns C1_OPT outputs 256 rows apart: a weight octet is read once (broadcast) for all of them
  float acc[C1_OPT];
  const float bv = bias != nullptr ? bias[0] : 0.f;
#pragma unroll
  for (int o = 0; o < C1_OPT; ++o) acc[o] = bv;
  for (int j = 0; j < ntaps; ++j) {
    const float* wj = ws + j * Cin;
    for (int u = 0; u < upr; ++u) {
      const float4 w0 = *reinterpret_cast<const float4*>(wj + u * 8), w1 = *reinterpret_cast<const float4*>(wj + u * 8 + 4);
#pragma unroll
      for (int o = 0; o < C1_OPT; ++o) {
        const int r = threadIdx.x + o * C1_THREADS + j * dil;
        const uint4 v = reinterpret_cast<const uint4*>(tile)[r * upr + ((u + r) % upr)];
        float a[8];
        if (BF16) {
          a[0] = __uint_as_float(v.x << 16); a[1] = __uint_as_float(v.x & 0xFFFF0000u);
          a[2] = __uint_as_float(v.y << 16); a[3] = __uint_as_float(v.y & 0xFFFF0000u);
          a[4] = __uint_as_float(v.z << 16); a[5] = __uint_as_float(v.z & 0xFFFF0000u);
          a[6] = __uint_as_float(v.w << 16); a[7] = __uint_as_float(v.w & 0xFFFF0000u);
        } else {
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
          const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
          a[0] = f0.x; a[1] = f0.y; a[2] = f1.x; a[3] = f1.y; a[4] = f2.x; a[5] = f2.y; a[6] = f3.x; a[7] = f3.y;
        }
        acc[o] += a[0] * w0.x + a[1] * w0.y + a[2] * w0.z + a[3] * w0.w + a[4] * w1.x + a[5] * w1.y + a[6] * w1.z + a[7] * w1.w;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < C1_OPT; ++o) {
    const int t = t0 + threadIdx.x + o * C1_THREADS;
    if (t >= T) continue;
    float v = acc[o] * out_scale;
    if (lens != nullptr && t >= lens[b]) v = 0.f;
    const long long row = (long long)b * T + t;
    if (y_raw) stany(y_raw, row * y_raw_ld, v, y_raw_dtype);
    if (y_act) {
      const float a = apply_act(v, act, slope);
      if (y_act_dtype == AS_PCM16)   // libsndfile's float -> PCM_16: lrint(32767 * x) (test.py:119), saturated
        reinterpret_cast<int16_t*>(y_act)[row * y_act_ld] = (int16_t)max(-32768, min(32767, __float2int_rn(a * 32767.f)));
      else
        stany(y_act, row * y_act_ld, a, y_act_dtype);
    }
  }
}

// eligible: 1-D, one output channel, 16-bit input with Cin % 8 == 0 and Cin <= 64, equally spaced taps, no residuals
bool conv_cout1_eligible(const as_conv_params* p) {
  if (p->Cout != 1 || p->F != 1 || p->Fo != 1 || p->To != p->T || p->res1 || p->res2 || p->stats) return false;
  if (p->y_raw && p->y_raw_dtype == AS_PCM16) return false;    // PCM is an activated-output format only
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
  static bool attr = false;
  if (!attr) {
    ASB_CUDA(cudaFuncSetAttribute(conv_cout1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ASB_CUDA(cudaFuncSetAttribute(conv_cout1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  dim3 grid((unsigned)((p->T + C1_ROWS - 1) / C1_ROWS), (unsigned)p->B);
  const long long w_tap_stride = (long long)p->CoutP * p->CinP;    // row 0 (the only output channel) of every tap
  if (p->x_dtype == AS_BF16)
    ASB_CUDA(launch_k(conv_cout1_kernel<true>, grid, C1_THREADS, smem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
        reinterpret_cast<const uint16_t*>(p->w), w_tap_stride, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
        p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
  else
    ASB_CUDA(launch_k(conv_cout1_kernel<false>, grid, C1_THREADS, smem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
        reinterpret_cast<const uint16_t*>(p->w), w_tap_stride, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
        p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

}  // namespace asb
