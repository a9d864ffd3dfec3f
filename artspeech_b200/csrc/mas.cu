// MAS maximum_path for sm_100a.
//
// Replaces S_monotonic_align.py:5-47 (maximum_path1), :50-95 (maximum_path2) and the Triton
// kernel S_monotonic_align_Triton.py:7-71 with ONE kernel.
//
// One CTA (4 warps) per batch item.  Warp 0 owns the dynamic programme: lane l keeps R
// consecutive rows (x = l*R .. l*R+R-1) of the running column in registers, so the
// "previous row, previous column" operand is a register for R-1 of its rows and one
// __shfl_up_sync for the first.  Column values arrive through a double-buffered, transposed
// shared-memory tile filled with cp.async by all 4 warps (global reads coalesced along Ty,
// shared reads conflict-free because R and the pitch are odd).  Instead of writing the
// cumulative matrix back (what the reference and the Triton kernel do), the forward pass keeps
// ONE direction bit per cell (shared memory, or the caller's workspace when Tx*Ty bits exceed
// it).  The back-track reads a 32x32 bit block per 32 columns, transposes it with 32 ballots
// and walks it with warp-uniform integer ops.  Warps 1-3 zero-fill `path` in the shadow of the
// forward pass.  `value` is never written (the Triton kernel mutates it).
//
// Arithmetic is exactly the reference's: q = value + (same > diag ? same : diag) in fp32, one
// add and one compare per cell, no re-association -> bit-identical cumulative values, hence
// bit-identical paths.
#include "common.cuh"

namespace asb {

constexpr int MAS_THREADS = 128;
constexpr float MAS_NEG = -1e32f;

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n"); }

template <int R>
__global__ void __launch_bounds__(MAS_THREADS)
mas_kernel(const float* __restrict__ value, const int32_t* __restrict__ x_len,
           const int32_t* __restrict__ y_len, float* __restrict__ path, int Tx, int Ty,
           int tie_move, uint32_t* __restrict__ dir_ws, int dir_in_smem, int cw) {
  constexpr int ROWS = 32 * R;     // rows covered by warp 0
  constexpr int PX = ROWS + 1;     // odd pitch of the transposed [col][row] tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                 // [2][cw][PX]
  uint32_t* dir_s = reinterpret_cast<uint32_t*>(tile + 2 * cw * PX);  // [ROWS][W] if in smem

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  int xl = x_len[b], yl = y_len[b];
  xl = min(max(xl, 0), Tx);
  yl = min(max(yl, 0), Ty);
  const int W = ((Ty + 31) >> 5) | 1;  // direction words per row (odd pitch: no bank conflicts)
  const float* vb = value + (size_t)b * Tx * Ty;
  float* pb = path + (size_t)b * Tx * Ty;
  uint32_t* dirs = dir_in_smem ? dir_s : (dir_ws + (size_t)b * ROWS * W);

  const int nwords = (yl + 31) >> 5;     // direction words actually used per row
  const int nsub = (yl + cw - 1) / cw;   // column sub-chunks of cw columns (cw in {32,16,8})

  // ---- zero-fill bookkeeping for warps 1..3 (whole item, spread over the sub-chunks) ----
  const size_t total = (size_t)Tx * Ty;
  const bool vec_ok = ((total & 3) == 0) && ((reinterpret_cast<uintptr_t>(pb) & 15) == 0);
  const size_t nz = vec_ok ? (total >> 2) : total;  // units to write
  const int zsteps = max(nsub, 1);
  const size_t zper = (nz + zsteps - 1) / zsteps;

  auto zero_slice = [&](int step) {
    if (warp == 0) return;
    size_t lo = (size_t)step * zper, hi = min(nz, lo + zper);
    if (vec_ok) {
      float4* p4 = reinterpret_cast<float4*>(pb);
      for (size_t i = lo + (tid - 32); i < hi; i += MAS_THREADS - 32) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (size_t i = lo + (tid - 32); i < hi; i += MAS_THREADS - 32) pb[i] = 0.f;
    }
  };

  auto load_sub = [&](int sc) {
    // rows [0, xl) x cols [cw*sc, cw*sc+cw) -> tile[sc&1][col][row]
    float* dst = tile + (sc & 1) * cw * PX;
    const int n = xl * cw;
    for (int i = tid; i < n; i += MAS_THREADS) {
      int col = i % cw, row = i / cw;
      int y = sc * cw + col;
      if (y < yl) cp_async4(smem_u32(dst + col * PX + row), vb + (size_t)row * Ty + y);
    }
    cp_async_commit();
  };

  if (nsub == 0 || xl == 0) {
    // empty item (x_len == 0 or y_len == 0): the path is all zeros
    for (size_t i = tid; i < total; i += MAS_THREADS) pb[i] = 0.f;
    return;
  }

  load_sub(0);
  cp_async_wait_all();
  __syncthreads();

  float q[R];
  uint32_t bits[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { q[r] = MAS_NEG; bits[r] = 0u; }

  for (int sc = 0; sc < nsub; ++sc) {
    if (sc + 1 < nsub) load_sub(sc + 1);
    zero_slice(sc);
    if (warp == 0) {
      const float* src = tile + (sc & 1) * cw * PX + lane * R;
      const int ncol = min(cw, yl - sc * cw);
      for (int col = 0; col < ncol; ++col) {
        const float* vc = src + col * PX;
        const int y = sc * cw + col;
        if (y == 0) {
          // first column: only (0,0) is reachable (S_monotonic_align.py:23 / :68)
#pragma unroll
          for (int r = 0; r < R; ++r) q[r] = MAS_NEG;
          if (lane == 0) q[0] = vc[0];
        } else {
          float up = __shfl_up_sync(0xffffffffu, q[R - 1], 1);
          if (lane == 0) up = MAS_NEG;  // row -1 (S_monotonic_align.py:28 / :73)
          const int sh = y & 31;
#pragma unroll
          for (int r = R - 1; r >= 0; --r) {
            const float same = q[r];
            const float diag = (r == 0) ? up : q[r - 1];
            const bool take_same = same > diag;  // torch.where(prev1 > prev2, prev1, prev2)
            const float best = take_same ? same : diag;
            // back-track rule for the transition y -> y-1 evaluated at this row:
            //   stay mode (maximum_path2 :91 / Triton :37): move iff diag > same
            //   move mode (maximum_path1 :40): move iff !(same > diag)
            const bool mv = tie_move ? !take_same : (diag > same);
            bits[r] |= (mv ? 1u : 0u) << sh;
            q[r] = vc[r] + best;
          }
        }
        if ((y & 31) == 31 || y == yl - 1) {
          // row 0 can never move to row -1 (maximum_path2 :91 `idx != 0`)
          if (lane == 0) bits[0] = 0u;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            dirs[(size_t)(lane * R + r) * W + (y >> 5)] = bits[r];
            bits[r] = 0u;
          }
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();

  // ---- back-track (warp 0) ----
  if (warp != 0) return;
  int idx0 = xl - 1;  // row occupied at the current column
  for (int w = nwords - 1; w >= 0; --w) {
    const int hi = min(31, yl - 1 - (w << 5));
    const int row = idx0 - lane;
    uint32_t word = (row >= 0) ? dirs[(size_t)row * W + w] : 0u;
    uint32_t m[32];
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) m[cc] = __ballot_sync(0xffffffffu, (word >> cc) & 1u);
    int r = 0;
    int myrow = -1;
#pragma unroll
    for (int cc = 31; cc >= 0; --cc) {
      if (cc <= hi) {
        if (lane == cc) myrow = idx0 - r;
        const int y = (w << 5) + cc;
        if (y >= 1) r += (m[cc] >> r) & 1u;  // transition y -> y-1
      }
    }
    if (lane <= hi && myrow >= 0) pb[(size_t)myrow * Ty + (w << 5) + lane] = 1.0f;
    idx0 -= r;
  }
}

template <int R>
static int launch_mas(const float* value, const int32_t* x_len, const int32_t* y_len, float* path,
                      int B, int Tx, int Ty, int tie_mode, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  constexpr int ROWS = 32 * R;
  const int W = ((Ty + 31) / 32) | 1;
  const size_t smem_cap = 200 * 1024;
  int cw = 32;  // columns per staged sub-chunk: shrink until the double buffer fits
  while (cw > 8 && (size_t)2 * cw * (ROWS + 1) * sizeof(float) > smem_cap / 2) cw >>= 1;
  const size_t tile_bytes = (size_t)2 * cw * (ROWS + 1) * sizeof(float);
  const size_t dir_bytes = (size_t)ROWS * W * sizeof(uint32_t);
  int dir_in_smem = (tile_bytes + dir_bytes <= smem_cap) ? 1 : 0;
  size_t smem = tile_bytes + (dir_in_smem ? dir_bytes : 0);
  if (!dir_in_smem) {
    ASB_REQUIRE(ws != nullptr && ws_bytes >= dir_bytes * (size_t)B, AS_ERR_WORKSPACE,
                "as_mas_maximum_path: workspace too small (%zu < %zu)", ws_bytes,
                dir_bytes * (size_t)B);
  }
  ASB_REQUIRE(tile_bytes <= smem_cap, AS_ERR_SHAPE, "as_mas_maximum_path: Tx=%d too large", Tx);
  ASB_CUDA(cudaFuncSetAttribute(mas_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  mas_kernel<R><<<B, MAS_THREADS, smem, st>>>(value, x_len, y_len, path, Tx, Ty, tie_mode,
                                               reinterpret_cast<uint32_t*>(ws), dir_in_smem, cw);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

// rows per lane: odd (conflict-free shared reads) and one of the instantiated values
static int mas_rows_per_lane(int Tx) {
  static const int kR[] = {1, 3, 5, 7, 9, 13, 19, 25, 33};
  const int need = (Tx + 31) / 32;
  for (int r : kR) if (r >= need) return r;
  return -1;
}

}  // namespace asb

extern "C" size_t as_mas_workspace_bytes(int32_t B, int32_t Tx, int32_t Ty) {
  if (B <= 0 || Tx <= 0 || Ty <= 0) return 0;
  int R = asb::mas_rows_per_lane(Tx);
  if (R < 0) return 0;
  size_t W = (size_t)((Ty + 31) / 32) | 1;
  return (size_t)B * 32 * R * W * sizeof(uint32_t);
}

extern "C" int as_mas_maximum_path(const float* value, const int32_t* x_len, const int32_t* y_len,
                                   float* path, int32_t B, int32_t Tx, int32_t Ty,
                                   int32_t tie_mode, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  using namespace asb;
  if (B == 0 || Tx == 0 || Ty == 0) return AS_OK;
  ASB_REQUIRE(B > 0 && Tx > 0 && Ty > 0, AS_ERR_SHAPE, "as_mas_maximum_path: bad shape");
  ASB_REQUIRE(value && x_len && y_len && path, AS_ERR_SHAPE, "as_mas_maximum_path: null pointer");
  ASB_REQUIRE(tie_mode == 0 || tie_mode == 1, AS_ERR_SHAPE, "as_mas_maximum_path: tie_mode");
  int rc = check_arch();
  if (rc != AS_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int R = mas_rows_per_lane(Tx);
#define MAS_CASE(RR)                                                                          \
  if (R == RR)                                                                                \
    return launch_mas<RR>(value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace,          \
                          workspace_bytes, st);
  MAS_CASE(1) MAS_CASE(3) MAS_CASE(5) MAS_CASE(7) MAS_CASE(9) MAS_CASE(13) MAS_CASE(19)
  MAS_CASE(25) MAS_CASE(33)
#undef MAS_CASE
  set_error("as_mas_maximum_path: Tx=%d exceeds the supported maximum of 1056", Tx);
  return AS_ERR_SHAPE;
}
