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

constexpr float MAS_NEG = -1e32f;

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n"); }

// Multi-warp wavefront: the rows are split over NW warps (32*R consecutive rows per warp, R rows per
// lane).  Warp w can sweep a block of `cw` columns as soon as warp w-1 has swept the same block (it
// needs that warp's last row, one column back), so at pipeline step tau warp w processes block
// tau - w: NW blocks are in flight, one __syncthreads per step.  Inside a warp the diagonal operand
// is a register (R > 1) or one __shfl_up_sync; across warps it travels through a tiny shared ring
// (`bnd`) that the consumer pre-loads into registers per block.
template <int R>
__global__ void __launch_bounds__(256)
mas_kernel(const float* __restrict__ value, const int32_t* __restrict__ x_len,
           const int32_t* __restrict__ y_len, float* __restrict__ path, int Tx, int Ty,
           int tie_move, uint32_t* __restrict__ dir_ws, int dir_in_smem, int cw, int NW,
           int32_t* __restrict__ dur, int32_t* __restrict__ tok) {
  pdl_wait();   // programmatic dependent launch: everything above overlaps the previous kernel's tail
  constexpr int WR = 32 * R;       // rows per warp
  constexpr int PX = WR + 1;       // odd pitch of a warp's transposed [col][row] sub-tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ROWS = WR * NW;
  float* tile = reinterpret_cast<float*>(smem_raw);                     // [2][NW][cw][PX]
  float* bnd = tile + 2 * NW * cw * PX;                                 // [NW][2][32] last-row values
  uint32_t* dir_s = reinterpret_cast<uint32_t*>(bnd + NW * 2 * 32);     // [ROWS][W] if in smem

  const int b = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  int xl = x_len[b], yl = y_len[b];
  xl = min(max(xl, 0), Tx);
  yl = min(max(yl, 0), Ty);
  const int W = ((Ty + 31) >> 5) | 1;  // direction words per row (odd pitch: no bank conflicts)
  const float* vb = value + (size_t)b * Tx * Ty;
  float* pb = path + (size_t)b * Tx * Ty;
  uint32_t* dirs = dir_in_smem ? dir_s : (dir_ws + (size_t)b * ROWS * W);

  const int nwords = (yl + 31) >> 5;
  const int nblk = (yl + cw - 1) / cw;          // column blocks
  const int nsteps = nblk > 0 ? nblk + NW - 1 : 0;

  const size_t total = (size_t)Tx * Ty;
  // optional alignment outputs (as_mas_align): durations = row sums of the path, tok = row of every column
  if (dur) for (int i = tid; i < Tx; i += nthr) dur[(size_t)b * Tx + i] = 0;
  if (tok) for (int i = tid; i < Ty; i += nthr) tok[(size_t)b * Ty + i] = -1;
  if (nblk == 0 || xl == 0) {
    // empty item (x_len == 0 or y_len == 0): the path is all zeros
    for (size_t i = tid; i < total; i += nthr) pb[i] = 0.f;
    return;
  }
  // zero-fill of `path`, spread over the pipeline steps (all threads; pure stores)
  const bool vec_ok = ((total & 3) == 0) && ((reinterpret_cast<uintptr_t>(pb) & 15) == 0);
  const size_t nz = vec_ok ? (total >> 2) : total;
  const size_t zper = (nz + nsteps - 1) / nsteps;
  auto zero_slice = [&](int step) {
    size_t lo = (size_t)step * zper, hi = min(nz, lo + zper);
    if (vec_ok) {
      float4* p4 = reinterpret_cast<float4*>(pb);
      for (size_t i = lo + tid; i < hi; i += nthr) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (size_t i = lo + tid; i < hi; i += nthr) pb[i] = 0.f;
    }
  };
  // stage the sub-tiles every warp needs at pipeline step `tau` (block tau - w for warp slot w)
  auto load_step = [&](int tau) {
    float* dst0 = tile + (size_t)(tau & 1) * NW * cw * PX;
    for (int w = 0; w < NW; ++w) {
      const int blk = tau - w;
      if (blk < 0 || blk >= nblk) continue;
      const int r0 = w * WR;
      const int nrow = min(WR, xl - r0);
      if (nrow <= 0) continue;
      float* dst = dst0 + (size_t)w * cw * PX;
      const int n = nrow * cw;
      for (int i = tid; i < n; i += nthr) {
        int col, row;
        if (cw == 32) { col = i & 31; row = i >> 5; } else { col = i % cw; row = i / cw; }
        const int y = blk * cw + col;
        if (y < yl) cp_async4(smem_u32(dst + col * PX + row), vb + (size_t)(r0 + row) * Ty + y);
      }
    }
    cp_async_commit();
  };

  for (int i = tid; i < NW * 2 * 32; i += nthr) bnd[i] = MAS_NEG;
  load_step(0);
  cp_async_wait_all();
  __syncthreads();

  float q[R];
  uint32_t bits[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { q[r] = MAS_NEG; bits[r] = 0u; }
  float carry = MAS_NEG;                       // boundary row value at the last column of my previous block
  const int row0 = (warp * 32 + lane) * R;     // first row of this lane
  const bool warp_active = warp * WR < xl;

  for (int tau = 0; tau < nsteps; ++tau) {
    if (tau + 1 < nsteps) load_step(tau + 1);
    zero_slice(tau);
    const int blk = tau - warp;
    if (warp_active && blk >= 0 && blk < nblk) {
      const float* src = tile + ((size_t)(tau & 1) * NW + warp) * cw * PX + lane * R;
      // boundary operands from the previous warp for this block's columns (lane c holds column c)
      float bv = MAS_NEG;
      if (warp > 0 && lane < cw) bv = bnd[((warp - 1) * 2 + (blk & 1)) * 32 + lane];
      float* my_bnd = bnd + (warp * 2 + (blk & 1)) * 32;
      const int ncol = min(cw, yl - blk * cw);
      if (cw == 32 && ncol == 32 && blk > 0) {
        // fast path (ncu: the generic loop below spends ~60 dependent instructions per column with one
        // or two warps per scheduler): a full 32-column block, fully unrolled -> static shuffle lanes,
        // static bit positions, immediate shared-memory offsets; exactly one direction word per row.
        const float up0 = (warp == 0) ? MAS_NEG : carry;
#pragma unroll
        for (int col = 0; col < 32; ++col) {
          const float* vc = src + col * PX;
          float up = __shfl_up_sync(0xffffffffu, q[R - 1], 1);
          const float prevw = __shfl_sync(0xffffffffu, bv, (col + 31) & 31);
          if (lane == 0) up = (col == 0) ? up0 : ((warp == 0) ? MAS_NEG : prevw);
#pragma unroll
          for (int r = R - 1; r >= 0; --r) {
            const float same = q[r];
            const float diag = (r == 0) ? up : q[r - 1];
            const bool take_same = same > diag;
            const float best = take_same ? same : diag;
            const bool mv = tie_move ? !take_same : (diag > same);
            bits[r] |= (mv ? 1u : 0u) << col;
            q[r] = vc[r] + best;
          }
          if (lane == 31) my_bnd[col] = q[R - 1];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          dirs[(size_t)(row0 + r) * W + blk] = (row0 + r == 0) ? 0u : bits[r];
          bits[r] = 0u;
        }
      } else
      for (int col = 0; col < ncol; ++col) {
        const float* vc = src + col * PX;
        const int y = blk * cw + col;
        if (y == 0) {
          // first column: only (0,0) is reachable (S_monotonic_align.py:23 / :68)
#pragma unroll
          for (int r = 0; r < R; ++r) q[r] = MAS_NEG;
          if (row0 == 0) q[0] = vc[0];
        } else {
          float up = __shfl_up_sync(0xffffffffu, q[R - 1], 1);
          // lane 0: row above belongs to the previous warp (column y-1): this block's column col-1,
          // or the carried last column of the previous block; warp 0 has row -1 (:28 / :73)
          const float prevw = __shfl_sync(0xffffffffu, bv, (col + 31) & 31);
          if (lane == 0) up = (warp == 0) ? MAS_NEG : (col == 0 ? carry : prevw);
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
        if (lane == 31) my_bnd[col] = q[R - 1];
        if ((y & 31) == 31 || y == yl - 1) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            // row 0 can never move to row -1 (maximum_path2 :91 `idx != 0`)
            dirs[(size_t)(row0 + r) * W + (y >> 5)] = (row0 + r == 0) ? 0u : bits[r];
            bits[r] = 0u;
          }
        }
      }
      // carry for my next block: the previous warp's value at the last column of this block
      carry = __shfl_sync(0xffffffffu, bv, (cw - 1) & 31);
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
    if (lane <= hi && myrow >= 0) {
      pb[(size_t)myrow * Ty + (w << 5) + lane] = 1.0f;
      if (tok) tok[(size_t)b * Ty + (w << 5) + lane] = myrow;
      if (dur) atomicAdd(&dur[(size_t)b * Tx + myrow], 1);
    }
    idx0 -= r;
  }
}

struct MasPlan { int R, NW, cw, W, dir_in_smem; size_t smem, dir_bytes_per_item; };

static bool mas_plan(int Tx, int Ty, MasPlan& p) {
  const int nwarp_rows = (Tx + 31) / 32;
  p.NW = nwarp_rows < 8 ? (nwarp_rows < 1 ? 1 : nwarp_rows) : 8;
  int need = (Tx + 32 * p.NW - 1) / (32 * p.NW);
  static const int kR[] = {1, 3, 5};
  p.R = -1;
  for (int r : kR) if (r >= need) { p.R = r; break; }
  if (p.R < 0) return false;
  const int ROWS = 32 * p.R * p.NW;
  p.W = ((Ty + 31) / 32) | 1;
  const size_t cap = 200 * 1024;
  p.cw = 32;
  auto tile_bytes = [&](int cw) { return (size_t)2 * p.NW * cw * (32 * p.R + 1) * sizeof(float); };
  while (p.cw > 8 && tile_bytes(p.cw) > cap / 2) p.cw >>= 1;
  const size_t fixed = tile_bytes(p.cw) + (size_t)p.NW * 2 * 32 * sizeof(float);
  p.dir_bytes_per_item = (size_t)ROWS * p.W * sizeof(uint32_t);
  p.dir_in_smem = (fixed + p.dir_bytes_per_item <= cap) ? 1 : 0;
  p.smem = fixed + (p.dir_in_smem ? p.dir_bytes_per_item : 0);
  return fixed <= cap;
}

template <int R>
static int launch_mas(const MasPlan& p, const float* value, const int32_t* x_len, const int32_t* y_len,
                      float* path, int B, int Tx, int Ty, int tie_mode, void* ws, cudaStream_t st,
                      int32_t* dur = nullptr, int32_t* tok = nullptr) {
  ASB_CUDA(cudaFuncSetAttribute(mas_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  ASB_CUDA(launch_k(mas_kernel<R>, B, 32 * p.NW, p.smem, st, value, x_len, y_len, path, Tx, Ty, tie_mode,
                                              reinterpret_cast<uint32_t*>(ws), p.dir_in_smem, p.cw, p.NW, dur, tok));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

}  // namespace asb

extern "C" size_t as_mas_workspace_bytes(int32_t B, int32_t Tx, int32_t Ty) {
  if (B <= 0 || Tx <= 0 || Ty <= 0) return 0;
  asb::MasPlan p;
  if (!asb::mas_plan(Tx, Ty, p)) return 0;
  return p.dir_in_smem ? 0 : (size_t)B * p.dir_bytes_per_item;
}

static int mas_run(const char* who, const float* value, const int32_t* x_len, const int32_t* y_len, float* path, int32_t B,
                   int32_t Tx, int32_t Ty, int32_t tie_mode, void* workspace, size_t workspace_bytes, void* stream,
                   int32_t* dur, int32_t* tok) {
  using namespace asb;
  if (B == 0 || Tx == 0 || Ty == 0) return AS_OK;
  ASB_REQUIRE(B > 0 && Tx > 0 && Ty > 0, AS_ERR_SHAPE, "%s: bad shape", who);
  ASB_REQUIRE(value && x_len && y_len && path, AS_ERR_SHAPE, "%s: null pointer", who);
  ASB_REQUIRE(tie_mode == 0 || tie_mode == 1, AS_ERR_SHAPE, "%s: tie_mode", who);
  int rc = check_arch();
  if (rc != AS_OK) return rc;
  MasPlan p;
  ASB_REQUIRE(mas_plan(Tx, Ty, p), AS_ERR_SHAPE, "%s: Tx=%d exceeds the supported maximum of 1280", who, Tx);
  if (!p.dir_in_smem) {
    ASB_REQUIRE(workspace != nullptr && workspace_bytes >= p.dir_bytes_per_item * (size_t)B, AS_ERR_WORKSPACE,
                "%s: workspace too small (%zu < %zu)", who, workspace_bytes, p.dir_bytes_per_item * (size_t)B);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.R == 1) return launch_mas<1>(p, value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace, st, dur, tok);
  if (p.R == 3) return launch_mas<3>(p, value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace, st, dur, tok);
  return launch_mas<5>(p, value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace, st, dur, tok);
}

extern "C" int as_mas_maximum_path(const float* value, const int32_t* x_len, const int32_t* y_len,
                                   float* path, int32_t B, int32_t Tx, int32_t Ty,
                                   int32_t tie_mode, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  return mas_run("as_mas_maximum_path", value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace, workspace_bytes, stream,
                 nullptr, nullptr);
}

extern "C" int as_mas_align(const float* value, const int32_t* x_len, const int32_t* y_len, float* path,
                            int32_t* durations, int32_t* token_of_frame, int32_t B, int32_t Tx, int32_t Ty,
                            int32_t tie_mode, void* workspace, size_t workspace_bytes, void* stream) {
  if (B > 0 && Tx > 0 && Ty > 0)
    ASB_REQUIRE(durations && token_of_frame, AS_ERR_SHAPE, "as_mas_align: null pointer");
  return mas_run("as_mas_align", value, x_len, y_len, path, B, Tx, Ty, tie_mode, workspace, workspace_bytes, stream,
                 durations, token_of_frame);
}
