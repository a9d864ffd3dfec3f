// Small device-side glue of the serving path: everything the reference does on the host between the duration
// predictor and the length regulator (models.py:361-366: round, clamp, cumulative frame counts) stays on the
// device, so the whole pass is capturable in a CUDA graph with durations as a graph INPUT.
#include "common.cuh"

namespace asb {

// One warp per utterance: dur[b,t] = max(1, rint(pred[b,t])) for t < lens[b], else 0 (torch.round is
// round-half-to-even = rintf in the default rounding mode); sum[b] = total half-rate frames of the utterance.
__global__ void round_durations_kernel(const float* __restrict__ pred, long long pred_ld, int B, int Tt,
                                       const int* __restrict__ lens, int* __restrict__ dur, int* __restrict__ sum) {
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = lens ? min(max(lens[b], 0), Tt) : Tt;
  int acc = 0;
  for (int t = lane; t < Tt; t += 32) {
    int d = 0;
    if (t < n) {
      const float r = rintf(pred[(long long)b * pred_ld + t]);
      d = r >= 1.f ? (r > 2147483000.f ? 2147483000 : (int)r) : 1;     // clamp(min=1); NaN -> 1
    }
    dur[(long long)b * Tt + t] = d;
    acc += d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && sum) sum[b] = acc;
}

}  // namespace asb

using namespace asb;

extern "C" int as_round_durations(const float* pred, int64_t pred_ld, int32_t B, int32_t Tt, const int32_t* lens,
                                  int32_t* dur, int32_t* sum, void* stream) {
  if (B == 0) return AS_OK;
  ASB_REQUIRE(pred && dur && B > 0 && Tt >= 0 && pred_ld >= Tt, AS_ERR_SHAPE, "as_round_durations: bad argument");
  if (int rc = check_arch()) return rc;
  const int warps = 4;
  ASB_CUDA(launch_k(round_durations_kernel, dim3((B + warps - 1) / warps), dim3(32 * warps), 0,
                    static_cast<cudaStream_t>(stream), pred, (long long)pred_ld, B, Tt, lens, dur, sum));
  return AS_OK;
}
