// Shared helpers for the sm_100a kernels behind the C ABI (include/artspeech_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/artspeech_b200.h"

namespace asb {

// ---- error plumbing (per-thread last-error string, no exceptions across the ABI) ----
void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);   // returns AS_OK or AS_ERR_CUDA
int  check_arch();                                  // AS_OK iff current device is sm_100

#define ASB_CUDA(call)                                                      \
  do {                                                                      \
    int _rc = ::asb::check_cuda((call), #call);                             \
    if (_rc != AS_OK) return _rc;                                           \
  } while (0)

#define ASB_REQUIRE(cond, code, ...)                                        \
  do {                                                                      \
    if (!(cond)) { ::asb::set_error(__VA_ARGS__); return (code); }          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 16-bit storage helpers (dtype: AS_F16 / AS_BF16)
__device__ __forceinline__ float ld16(const void* p, int64_t i, int dtype) {
  if (dtype == AS_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ uint16_t to16(float v, int dtype) {
  if (dtype == AS_F16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t u, int dtype) {
  if (dtype == AS_F16) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int dtype) {
  return uint32_t(to16(a, dtype)) | (uint32_t(to16(b, dtype)) << 16);
}
// generic element load: dtype in {AS_F16, AS_BF16, AS_F32}
__device__ __forceinline__ float ldany(const void* p, int64_t i, int dtype) {
  if (dtype == AS_F32) return reinterpret_cast<const float*>(p)[i];
  return ld16(p, i, dtype);
}
__device__ __forceinline__ void stany(void* p, int64_t i, float v, int dtype) {
  if (dtype == AS_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<uint16_t*>(p)[i] = to16(v, dtype);
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case AS_ACT_LRELU: return v > 0.f ? v : v * slope;
    case AS_ACT_RELU:  return fmaxf(v, 0.f);
    case AS_ACT_TANH:  return tanhf(v);
    case AS_ACT_SWISH: return v / (1.f + __expf(-v));
    case AS_ACT_ABS:   return fabsf(v);
    default:           return v;
  }
}

}  // namespace asb
