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

// ---------------------------------------------------------------------------------------------
// Tensor-core variant (Cin % 16 == 0, at most 8 taps).  The kernel above spends ~490 instructions per output, half
// of them unpacking 16-bit inputs for the seven taps that read each row: it is issue-bound at 1.1 TB/s.  Here the
// TAPS become the N dimension of an mma.sync m16n8k16: P[row][tap] = sum_c x[row][c] * w[tap][c] for every staged
// row (one ldmatrix + one MMA per 16 rows x 16 channels, no unpacking), P goes to shared memory in fp32 and an
// output is the diagonal sum y[t] = bias + sum_j P[t + j * dil][j]: ~25 instructions per output.
// ---------------------------------------------------------------------------------------------
constexpr int C1M_ROWS = 256;                            // output rows per CTA
constexpr int C1M_PP = 9;                                // pitch (floats) of a P row: conflict-free diagonal reads

template <bool BF16>
__global__ void __launch_bounds__(256)
conv_cout1_mma_kernel(const uint16_t* __restrict__ x, long long x_ld, int T, int Cin, const uint16_t* __restrict__ w,
                      long long w_tap_stride, int ntaps, int dt0, int dil, const float* __restrict__ bias, const int* __restrict__ lens,
                      float out_scale, int act, float slope, void* y_raw, int y_raw_dtype, long long y_raw_ld,
                      void* y_act, int y_act_dtype, long long y_act_ld) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  extern __shared__ __align__(16) unsigned char c1m_smem[];
  const int b = blockIdx.y, t0 = blockIdx.x * C1M_ROWS;
  const int halo = (ntaps - 1) * dil;
  const int nrows = (C1M_ROWS + halo + 15) & ~15;         // staged rows, a multiple of the MMA's 16
  const int pitch = Cin * 2 + 16;                         // bytes per staged row: +16 keeps ldmatrix conflict-free
  unsigned char* xs = c1m_smem;                           // [nrows][pitch]
  float* P = reinterpret_cast<float*>(c1m_smem + (size_t)nrows * pitch);   // [nrows][C1M_PP]
  const int upr = Cin / 8;                                // 16-byte units per row
  const uint16_t* xb = x + (long long)b * T * x_ld;
  // 16-byte cp.async with zero fill for rows outside [0, T): every copy of the tile is in flight at once (with loads
  // through registers the ~8 dependent global loads per thread set the CTA's time: 143 us for the vocoder's conv_post)
  const uint32_t xs_base = smem_u32(xs);
  for (int i = threadIdx.x; i < nrows * upr; i += blockDim.x) {
    const int r = i / upr, u = i - r * upr;
    const int t = t0 + dt0 + r;
    const bool ok = t >= 0 && t < T;
    const void* src = ok ? static_cast<const void*>(reinterpret_cast<const uint4*>(xb + (long long)t * x_ld) + u) : static_cast<const void*>(xb);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs_base + (uint32_t)r * pitch + u * 16), "l"(src), "r"(ok ? 16u : 0u) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tid4 = lane & 3;
  // B fragments (weights): column n = tap gid, rows k = channels; taps >= ntaps are zero columns
  uint32_t bw[4][2];
  const int ksteps = Cin / 16;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    bw[ks][0] = bw[ks][1] = 0u;
    if (ks < ksteps && gid < ntaps) {
      const uint16_t* wr = w + (long long)gid * w_tap_stride + ks * 16 + tid4 * 2;
      bw[ks][0] = *reinterpret_cast<const uint32_t*>(wr);
      bw[ks][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t xs_u = xs_base;
  for (int mt = warp; mt < nrows / 16; mt += 8) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < ksteps) {
        uint32_t a[4];
        const uint32_t addr = xs_u + (uint32_t)(mt * 16 + (lane & 15)) * pitch + (uint32_t)(ks * 32 + (lane >> 4) * 16);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
        if (BF16)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                       : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bw[ks][0]), "r"(bw[ks][1]));
        else
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                       : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bw[ks][0]), "r"(bw[ks][1]));
      }
    }
    float* pr = P + (size_t)(mt * 16 + gid) * C1M_PP + tid4 * 2;
    pr[0] = d[0]; pr[1] = d[1];
    pr[8 * C1M_PP] = d[2]; pr[8 * C1M_PP + 1] = d[3];
  }
  __syncthreads();
  const float bv = bias != nullptr ? bias[0] : 0.f;
  const int len = lens != nullptr ? lens[b] : T;
  for (int o = threadIdx.x; o < C1M_ROWS; o += blockDim.x) {
    const int t = t0 + o;
    if (t >= T) break;
    float acc = bv;
    for (int j = 0; j < ntaps; ++j) acc += P[(size_t)(o + j * dil) * C1M_PP + j];
    float v = acc * out_scale;
    if (t >= len) v = 0.f;
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
  static const bool no_mma = getenv("ASB_COUT1_NO_MMA") != nullptr;
  if (!no_mma && (p->Cin % 16) == 0 && p->ntaps <= 8) {
    const int mrows = (C1M_ROWS + (p->ntaps - 1) * dil + 15) & ~15;
    const size_t msmem = (size_t)mrows * (p->Cin * 2 + 16) + (size_t)mrows * C1M_PP * sizeof(float);
    ASB_SMEM_OPT_IN(120 * 1024, conv_cout1_mma_kernel<true>);
    ASB_SMEM_OPT_IN(120 * 1024, conv_cout1_mma_kernel<false>);
    dim3 mgrid((unsigned)((p->T + C1M_ROWS - 1) / C1M_ROWS), (unsigned)p->B);
    const long long wts = (long long)p->CoutP * p->CinP;
    if (p->x_dtype == AS_BF16)
      ASB_CUDA(launch_k(conv_cout1_mma_kernel<true>, mgrid, 256, msmem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
          reinterpret_cast<const uint16_t*>(p->w), wts, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
          p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
    else
      ASB_CUDA(launch_k(conv_cout1_mma_kernel<false>, mgrid, 256, msmem, st, reinterpret_cast<const uint16_t*>(p->x), p->x_ld, p->T, p->Cin,
          reinterpret_cast<const uint16_t*>(p->w), wts, p->ntaps, p->tap_dt[0], dil, p->bias, p->lens, p->out_scale,
          p->act, p->slope, p->y_raw, p->y_raw_dtype, p->y_raw_ld, p->y_act, p->y_act_dtype, p->y_act_ld));
    ASB_CUDA(cudaGetLastError());
    return AS_OK;
  }
  const int nrows = C1_ROWS + (p->ntaps - 1) * dil;
  const size_t smem = (size_t)nrows * p->Cin * 2 + (size_t)p->ntaps * p->Cin * 4;
  ASB_SMEM_OPT_IN(160 * 1024, conv_cout1_kernel<true>);
  ASB_SMEM_OPT_IN(160 * 1024, conv_cout1_kernel<false>);
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
