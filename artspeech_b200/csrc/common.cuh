// Shared helpers for the sm_100a kernels behind the C ABI (include/artspeech_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

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

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is PER DEVICE, so the "already done"
// flag is a bit per device ordinal (a process may drive several GPUs, from several threads): one relaxed load on
// the hot path, the attribute call once per (kernel, device).  Usage: ASB_SMEM_OPT_IN(bytes, kernel<template args>).
#define ASB_SMEM_OPT_IN(bytes, ...)                                                                         \
  do {                                                                                                      \
    static std::atomic<unsigned long long> _asb_done[4];                                                    \
    int _asb_dev = 0;                                                                                       \
    ASB_CUDA(cudaGetDevice(&_asb_dev));                                                                     \
    const unsigned long long _asb_bit = 1ull << (_asb_dev & 63);                                            \
    std::atomic<unsigned long long>& _asb_w = _asb_done[(_asb_dev >> 6) & 3];                               \
    if (!(_asb_w.load(std::memory_order_acquire) & _asb_bit)) {                                             \
      ASB_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));    \
      _asb_w.fetch_or(_asb_bit, std::memory_order_release);                                                 \
    }                                                                                                       \
  } while (0)

// ---- programmatic dependent launch (PDL) ----
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization and executes
// griddepcontrol.wait before it touches data produced by the previous kernel of its stream: the launch
// latency and the kernel's own prologue then overlap the tail of the previous kernel (the synthesis step is
// a dependent chain of ~390 launches inside one CUDA graph).  The wait returns only once the previous grid
// has completed and flushed its writes, so data ordering is unchanged.  No kernel triggers its dependents
// early (griddepcontrol.launch_dependents): measured, the waiting CTAs then take SMs from the other
// pipelined batch and the step gets slower (16.3 -> 17.1 ms).  ASB_NO_PDL=1 falls back to ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = getenv("ASB_NO_PDL") == nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

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
__device__ __forceinline__ uint32_t pack16(float a, float b, int dtype) {   // one F2FP.PACK_AB: a -> low half, b -> high half
  if (dtype == AS_F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
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
