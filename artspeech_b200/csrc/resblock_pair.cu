// Fused HiFi-GAN ResBlock1 "pair" for sm_100a (Vocoder/vocoder.py:35-42):
//
//     xt = conv1_{k, dilation d}(leaky_relu(x)) ; xt = conv2_{k, 1}(leaky_relu(xt)) ; x = xt + x
//
// as ONE kernel per pair, for the vocoder stages with C <= 128 channels.  ncu on the layer-by-layer
// path (profiles/r01_vocoder_traffic_v11.txt) showed those stages moving 12 tensor-sized HBM
// transfers per pair (operand + residual in, raw + activated out, twice) and spending most of their
// time in the epilogue's load/store round trips.  Here the intermediate never leaves the SM:
//
//   * the residual stream is carried in its ACTIVATED form a = leaky_relu(x) only (one rounding to
//     16 bits instead of two).  TMA loads the tile of `a` with its halo straight into the swizzled
//     K-major layout the UMMA descriptor reads; the raw residual is recovered exactly where it is
//     needed as x = a < 0 ? a / slope : a;
//   * stage 1: D1[256 rows, C] = sum over taps of A(shifted by j*d rows) * W1_j^T  (tcgen05.mma, fp32
//     accumulators in TMEM, a tap is a row offset of the descriptor's start address);
//   * epilogue 1 (warpgroup 0): TMEM -> +bias -> LeakyReLU -> zero outside the sequence (conv2's
//     padding) -> 16-bit -> written IN PLACE over the A tile in the same swizzled layout; before
//     that the same warps seed stage 2's accumulator with residual + bias2 (tcgen05.st);
//   * stage 2: D2[256 rows, C] += sum over taps of T(shifted by j rows) * W2_j^T;
//   * epilogue 2 (warpgroup 1): TMEM -> (+ the other MRF branches) * scale -> mask -> activation ->
//     16-bit -> swizzled staging tile -> TMA store.  256 - (k-1) rows of every tile are valid.
//
// HBM traffic per pair: (1 + halo) reads + 1 write of the tensor instead of 12.  Weights stay
// resident in shared memory when they fit, otherwise they stream from L2 through an mbarrier ring.
// Warp roles (640 threads, one CTA per SM): 0 = activation-tile TMA producer, 1 = MMA issuer,
// 2 = weight TMA producer + TMEM allocator, 3 = second MMA issuer, 4-11 = stage-1 epilogue group (one warp per
// 32 rows of each M-tile), 12-19 = stage-2 epilogue group.  A lone warp per scheduler converts ~0.3
// instructions per clock: with 4 + 4 epilogue warps the kernel was epilogue-bound (~7 us per 256-row tile at
// C = 128, independent of k), so each group got a warp per (M-tile, TMEM lane quadrant).
#include "tc_util.cuh"

namespace asb {

constexpr int RP_THREADS = 640;

// Development aid (ASB_PAIR_DBG bit 5): CTA 0 records clock64() at the hand-off points of its first 32 tiles,
// trace[event * 32 + tile]; read back with as_debug_pair_trace().  Events: 0 producer got the buffer, 1/2 stage-1
// MMAs issue begin/end (M-tile 0), 3/4 stage-2, 5 group-1 has tile + D2 buffer, 6 seed done, 7 D1 ready, 8 group
// synced, 9 epilogue-1 math done, 10 group-1 signalled, 11 group-2 sees D2, 12 group-2 drained, 13 x tile landed.
__device__ unsigned long long g_pair_trace[20 * 32];
#define RP_TRACE(ev, tile_i)                                                                          \
  do {                                                                                                \
    if ((a.dbg & 32) && blockIdx.x == 0 && (tile_i) < 32) g_pair_trace[(ev) * 32 + (tile_i)] = clock64(); \
  } while (0)

struct PairArgs {
  int B, L, tiles_per_item, total_tiles;
  int k, dil, p1, p2, R_out, HA, HB, tail_rows;
  int NX, ND, NSTG, SW, stream1, stream2, pipelined;
  int pair;  // 1: cluster of two CTAs, tcgen05.mma.cta_group::2 (each CTA holds half of every weight block)
  int NT;    // stage-2 operand buffers: 0 = written in place over the input tile, 1..2 = separate buffers (the
             // input tile is then released as soon as stage 1 and the residual read are done)
  int dbg;   // timing experiments only (ASB_PAIR_DBG; results are wrong): 1 no MMAs, 2 no stores, 4 no reloads, 8 no epilogue-1 math, 16 no seed;
             // 64 no epilogue fast paths, 128 no tile skipping (results stay correct)
  uint32_t idesc;
  const float* b1;
  const float* b2;
  const void* res2; long long res2_ld;
  const void* res3; long long res3_ld;
  float slope, inv_slope, out_scale, out_slope_eff;
  const int* lens;
  void* y; long long y_ld;    // raw output pointer: skipped tiles are zero-filled with plain stores
};
struct PairMaps { CUtensorMap x, w1, w2, y, y_tail; };

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
template <bool BF16>
__device__ __forceinline__ void unpack2(uint32_t u, float& a, float& b) {
  if (BF16) { a = __uint_as_float(u << 16); b = __uint_as_float(u & 0xFFFF0000u); return; }
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
  a = f.x; b = f.y;
}

// Ring position + phase parity, advanced incrementally: every role of the kernel walks its buffers once per tile, and a
// runtime `i % n` / `i / n` costs ~100 dependent cycles each on what is a single-thread critical path (measured with
// the hand-off trace: four of them per stage put ~450 idle cycles between two MMA batches).
struct RingPos {
  int idx, n;
  uint32_t ph;
  __device__ __forceinline__ RingPos(int n_) : idx(0), n(n_), ph(0) {}
  __device__ __forceinline__ void next() { if (++idx == n) { idx = 0; ph ^= 1u; } }
};
// (item, tile-in-item) of the CTA's i-th tile, advanced by gridDim.x tiles per step
struct TilePos {
  int b, tt, step, per_item;
  __device__ __forceinline__ TilePos(int first, int step_, int per_item_) : step(step_), per_item(per_item_) {
    b = first / per_item_; tt = first - b * per_item_;
  }
  __device__ __forceinline__ void next() { tt += step; while (tt >= per_item) { tt -= per_item; ++b; } }
};

// shared-memory plan, computed identically on host and device
struct PairSmem {
  uint32_t xb, x_off, tb, t_off, wres_off, ring_off, stg_off, bias_off, bar_off, total;
};
constexpr uint32_t RP_NBARS = 96;
__host__ __device__ inline PairSmem pair_smem(int C, int HA, int k, int NX, int NT, int NSTG, int SW, int stream1, int stream2, int pair = 0) {
  const uint32_t BKC = C >= 64 ? 64 : 32, RB = BKC * 2, KCH = C / BKC, WBLK = (uint32_t)(pair ? C / 2 : C) * RB;
  PairSmem s;
  s.xb = ((KCH * (uint32_t)HA * RB) + 1023u) & ~1023u;
  s.x_off = 0;
  s.tb = KCH * 256u * RB;                       // 256 rows of conv1 output per tile (multiple of 1024 bytes)
  s.t_off = s.x_off + NX * s.xb;
  s.wres_off = s.t_off + NT * s.tb;
  const uint32_t nres = (stream1 ? 0 : k * KCH) + (stream2 ? 0 : k * KCH);
  s.ring_off = s.wres_off + nres * WBLK;
  s.stg_off = s.ring_off + SW * WBLK;
  s.bias_off = s.stg_off + 8u * NSTG * 32u * RB;
  s.bar_off = s.bias_off + 2u * C * 4u;
  s.total = s.bar_off + 8u * RP_NBARS;
  return s;
}

// PAIR: the kernel runs as clusters of two CTAs (one TPC).  Each CTA keeps its own tiles, buffers and epilogue groups;
// the LEADER's two issuer warps drive both tensor cores with tcgen05.mma.cta_group::2 (UMMA M = 256: M-tile mt of both
// CTAs in one instruction), and every weight block is split between the two CTAs' shared memories (C / 2 rows each),
// so the operand bytes a CTA fetches per instruction drop from A + W to A + W / 2 (the pair kernels' MMA phases run at
// the shared-memory operand limit).  Barrier topology: barriers the issuers WAIT on (x_full, d1_empty,
// t_full, d2_init, weight-full) live in the leader and collect both CTAs' arrivals (TMA completions with
// .cta_group::2, remote mbarrier arrives); barriers the issuers SIGNAL (d1_full, d2_full, x_empty, t_empty,
// weight-empty) exist in both CTAs and are hit by one multicast tcgen05.commit.  x_full is relayed to both CTAs'
// epilogue groups by a spare thread of the leader (x_seen).
template <int C, bool BF16, bool PAIR>
__device__ __forceinline__ void resblock_pair_body(const PairMaps& maps, const PairArgs& a) {
  constexpr int BKC = C >= 64 ? 64 : 32;       // channels per K chunk (one swizzled row)
  constexpr uint32_t RB = BKC * 2;             // bytes per row of a chunk
  constexpr int KCH = C / BKC;
  constexpr int WROWS = PAIR ? C / 2 : C;      // weight rows (output channels) this CTA holds of every block
  constexpr uint32_t WBLK = (uint32_t)WROWS * RB;  // one (tap, chunk) weight block (this CTA's share)
  constexpr int KS = BKC / 16;                 // MMAs (K = 16) per chunk
  constexpr int UPC = BKC / 8;                 // 16-byte units per chunk row
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const PairSmem sp = pair_smem(C, a.HA, a.k, a.NX, a.NT, a.NSTG, a.SW, a.stream1, a.stream2, a.pair);
  const int NX = a.NX, ND = a.ND, SW = a.SW, NT = a.NT;
  const uint32_t t_chunk_bytes = 256u * RB;
  const uint32_t chunk_bytes = (uint32_t)a.HA * RB;
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (base + sp.bias_off - smem_u32(smem_dyn)));
  const uint32_t bars = base + sp.bar_off;
  // barrier map (8 bytes each)
  auto x_full = [&](int i) { return bars + 8u * (64 + i); };     // [8]
  auto x_empty = [&](int i) { return bars + 8u * (72 + i); };    // [8]
  auto t_full = [&](int i) { return bars + 8u * (8 + i); };      // [4]
  auto t_empty = [&](int i) { return bars + 8u * (80 + i); };    // [4]
  auto x_seen = [&](int i) { return bars + 8u * (84 + i); };     // [8] pair mode: "input tile i landed in both CTAs"
  auto d1_full = [&](int i) { return bars + 8u * (12 + i); };    // [2]
  auto d1_empty = [&](int i) { return bars + 8u * (14 + i); };   // [2]
  auto d2_init = [&](int i) { return bars + 8u * (16 + i); };    // [2]
  auto d2_full = [&](int i) { return bars + 8u * (18 + i); };    // [2]
  auto d2_empty = [&](int i) { return bars + 8u * (20 + i); };   // [2]
  const uint32_t wres_full = bars + 8u * 22;
  auto wr_full = [&](int i) { return bars + 8u * (24 + i); };    // [16]
  auto wr_empty = [&](int i) { return bars + 8u * (40 + i); };   // [16]
  const uint32_t tmem_slot = bars + 8u * 56;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // tile walk: a CTA (1-CTA mode) or a cluster (pair mode: the two CTAs take adjacent tiles) per schedule slot
  const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_per_slot = PAIR ? 2 : 1;
  const int first_tile = sched_id * tiles_per_slot + (int)rank;
  const int tile_step = sched_n * tiles_per_slot;
  const int n_slots = (a.total_tiles + tiles_per_slot - 1) / tiles_per_slot;
  // barriers in the leader (as shared::cluster addresses when this is the peer)
  auto on_leader = [&](uint32_t bar) { return (PAIR && !leader) ? mapa_u32(bar, 0) : bar; };
  auto arrive_leader = [&](uint32_t bar) {
    if (PAIR) mbar_arrive_cluster(mapa_u32(bar, 0));
    else mbar_arrive(bar);
  };
  auto arrive_leader_relaxed = [&](uint32_t bar) {     // TMEM-only hand-offs (no memory writes to publish)
    if (PAIR) mbar_arrive_cluster_relaxed(mapa_u32(bar, 0));
    else mbar_arrive(bar);
  };
  auto commit = [&](uint32_t bar) {
    if (PAIR) tc2_commit(bar);
    else tc_commit(bar);
  };
  // ---- tile skipping: a tile whose first output row lies beyond its utterance's length produces only zeros.  Every
  // role evaluates the same pure function of (tile, lens) and leaves such tiles out of the pipeline altogether (no
  // barrier, ring position or buffer is touched); the drain group zero-fills their rows.  In pair mode the two CTAs'
  // adjacent tiles are skipped together or not at all.  Ragged batches (predicted durations, mixed-length serving)
  // otherwise compute their padding: every kernel masks by length, none skipped.
  const bool can_skip = a.lens != nullptr && !(a.dbg & 128);
  auto len_of = [&](int b) { return b >= a.B ? 0 : (a.lens != nullptr ? min(__ldg(a.lens + b), a.L) : a.L); };
  auto dead = [&](int b, int tt) { return tt * a.R_out >= len_of(b); };          // own tile is all padding (or the odd pair's filler)
  auto skip_tile = [&](const TilePos& tp) {
    if (!can_skip) return false;
    if (!dead(tp.b, tp.tt)) return false;
    if (!PAIR) return true;
    int pb = tp.b, pt = tp.tt + (leader ? 1 : -1);                                // the partner's tile
    if (pt >= a.tiles_per_item) { pt = 0; ++pb; } else if (pt < 0) { pt = a.tiles_per_item - 1; --pb; }
    return dead(pb, pt);
  };
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) bias_s[i] = i < C ? a.b1[i] : a.b2[i - C];

  if (threadIdx.x == 0) {
    // separate stage-2 buffers: the input tile is released by the two stage-1 commits + the eight residual readers
    const uint32_t ng = PAIR ? 16u : 8u;   // epilogue-group warps that arrive on a leader barrier (both CTAs' in pair mode)
    for (int i = 0; i < 8; ++i) { mbar_init(x_full(i), 1); mbar_init(x_empty(i), a.NT ? 10 : 2); mbar_init(x_seen(i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(t_full(i), ng); mbar_init(t_empty(i), 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full(i), 2); mbar_init(d1_empty(i), ng); mbar_init(d2_init(i), ng);
      mbar_init(d2_full(i), 2); mbar_init(d2_empty(i), 8);
    }
    mbar_init(wres_full, 1);
    for (int i = 0; i < 16; ++i) { mbar_init(wr_full(i), 1); mbar_init(wr_empty(i), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.y) : "memory");
  }
  if (warp == 2 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w2) : "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync();     // both CTAs' barriers are initialised and their TMEM is allocated (also a CTA barrier)
  else __syncthreads();
  tc_fence_after();
  pdl_wait();   // barrier init / TMEM alloc / descriptor prefetch above overlap the previous kernel's tail
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // TMEM columns: D1[buf][mtile] at (buf*2 + mtile)*C, D2[buf][mtile] at (2*ND + buf*2 + mtile)*C
  auto d1_col = [&](int buf, int mt) { return tmem_base + (uint32_t)((buf * 2 + mt) * C); };
  auto d2_col = [&](int buf, int mt) { return tmem_base + (uint32_t)((2 * ND + buf * 2 + mt) * C); };

  const int n_local = (sched_id < n_slots) ? (n_slots - sched_id + sched_n - 1) / sched_n : 0;   // same in both CTAs of a pair
  const int nblk = a.k * KCH;                                     // weight blocks per conv
  const uint32_t w1_res = base + sp.wres_off;                     // resident conv1 blocks (if !stream1)
  const uint32_t w2_res = base + sp.wres_off + (a.stream1 ? 0u : (uint32_t)nblk * WBLK);
  const uint32_t ring = base + sp.ring_off;

  if (warp == 0) {
    if (PAIR && leader && lane == 1) {
      // ===== relay (pair mode): x_full lives in the leader only; tell both CTAs' epilogue groups that tile i landed.
      // A thread of its own because the cluster-scope release of the remote arrive stalls its issuer for ~1300 cycles.
      RingPos xr(NX);
      TilePos tp(first_tile, tile_step, a.tiles_per_item);
      for (int i = 0; i < n_local; ++i, tp.next()) {
        if (skip_tile(tp)) continue;
        mbar_wait(x_full(xr.idx), xr.ph);
        mbar_arrive(x_seen(xr.idx));
        mbar_arrive_cluster(mapa_u32(x_seen(xr.idx), 1));
        xr.next();
      }
    }
    if (lane == 0) {
      // ===== activation-tile producer =====
      RingPos xr(NX);
      TilePos tp(first_tile, tile_step, a.tiles_per_item);
      for (int i = 0; i < n_local; ++i, tp.next()) {
        if (skip_tile(tp)) continue;
        const int b = tp.b, tt = tp.tt;
        const int xb = xr.idx;
        const uint32_t ph = xr.ph;
        xr.next();
        mbar_wait(x_empty(xb), ph ^ 1u);
        RP_TRACE(0, i);
        if ((a.dbg & 4) && i >= NX) { if (leader) mbar_arrive(x_full(xb)); continue; }
        // pair mode: both CTAs' copies complete on the LEADER's barrier, which expects the bytes of both
        if (leader) mbar_expect_tx(x_full(xb), (uint32_t)(PAIR ? 2 : 1) * KCH * chunk_bytes);
        const uint32_t xfb = on_leader(x_full(xb));
        const int row0 = tt * a.R_out - a.p2 - a.p1;
        const uint32_t dst = base + sp.x_off + xb * sp.xb;
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          for (int h = 0; h < 2; ++h)
            if (PAIR) tma2_load_3d(dst + c * chunk_bytes + h * a.HB * RB, &maps.x, xfb, c * BKC, row0 + h * a.HB, b);
            else tma_load_3d(dst + c * chunk_bytes + h * a.HB * RB, &maps.x, xfb, c * BKC, row0 + h * a.HB, b);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===== weight producer: resident blocks once, streamed blocks through the ring every tile =====
      const int nres = (a.stream1 ? 0 : nblk) + (a.stream2 ? 0 : nblk);
      if (nres > 0) {
        if (leader) mbar_expect_tx(wres_full, (uint32_t)(PAIR ? 2 : 1) * nres * WBLK);
        const uint32_t wfb = on_leader(wres_full);
        auto wload = [&](uint32_t dst, const CUtensorMap* wm, uint32_t bar, int j) {   // this CTA's rows of block j
          if (PAIR) tma2_load_2d(dst, wm, bar, (j % KCH) * BKC, (j / KCH) * C + (int)rank * WROWS);
          else tma_load_2d(dst, wm, bar, (j % KCH) * BKC, (j / KCH) * C);
        };
        if (!a.stream1)
          for (int j = 0; j < nblk; ++j) wload(w1_res + j * WBLK, &maps.w1, wfb, j);
        if (!a.stream2)
          for (int j = 0; j < nblk; ++j) wload(w2_res + j * WBLK, &maps.w2, wfb, j);
      }
      if (a.stream1 || a.stream2) {
        // the ring is filled in the order the MMA issuers consume it: stage 1 and stage 2 of every tile, and in
        // pipelined mode S1(0), then S1(i+1) before S2(i)
        RingPos wr(SW);
        auto feed = [&](int conv) {
          if (!(conv == 0 ? a.stream1 : a.stream2)) return;
          const CUtensorMap* wm = conv == 0 ? &maps.w1 : &maps.w2;
          for (int j = 0; j < nblk; ++j, wr.next()) {
            const int slot = wr.idx;
            const uint32_t ph = wr.ph;
            mbar_wait(wr_empty(slot), ph ^ 1u);
            if (leader) mbar_expect_tx(wr_full(slot), (uint32_t)(PAIR ? 2 : 1) * WBLK);
            if (PAIR) tma2_load_2d(ring + slot * WBLK, wm, on_leader(wr_full(slot)), (j % KCH) * BKC, (j / KCH) * C + (int)rank * WROWS);
            else tma_load_2d(ring + slot * WBLK, wm, wr_full(slot), (j % KCH) * BKC, (j / KCH) * C);
          }
        };
        TilePos tp(first_tile, tile_step, a.tiles_per_item);
        bool pending = false;                  // a live tile whose stage 2 has not been fed yet (pipelined order)
        for (int i = 0; i < n_local; ++i, tp.next()) {
          if (skip_tile(tp)) continue;
          feed(0);
          if (a.pipelined) { if (pending) feed(1); pending = true; }
          else feed(1);
        }
        if (pending) feed(1);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===== MMA issuers: warp 1 owns M-tile 0, warp 3 owns M-tile 1 (separate accumulators, separate
    // schedulers).  At N = C <= 128 one thread cannot issue tcgen05.mma fast enough to keep the tensor
    // pipe busy (measured ~100 clk per instruction with a single issuer), so the issue stream is split
    // and kept branch-free: the warp stays converged, one elected lane issues. =====
    const int mt = warp == 1 ? 0 : 1;
    if (leader && elect_one()) {
      if (!a.stream1 || !a.stream2) mbar_wait(wres_full, 0);
      int slot = 0;            // weight ring position (streamed convolutions, in-order mode only)
      uint32_t slot_ph = 0;
      const uint64_t x_chunk_step = (uint64_t)(chunk_bytes >> 4), t_chunk_step = (uint64_t)(t_chunk_bytes >> 4);
      const uint32_t idesc = a.idesc;
      const int k = a.k;
      const bool do_mma = !(a.dbg & 1);
      // one convolution of one M-tile: k taps x KCH chunks x KS instructions into accumulator `tacc`
      auto run_conv = [&](uint32_t tacc, uint64_t da_tap, uint64_t tap_step, uint64_t chunk_step, uint32_t w_res, bool streamed, uint32_t acc) {
        if (!streamed) {
          uint64_t dw = make_smem_desc<BKC>(w_res);
          for (int tap = 0; tap < k; ++tap, da_tap += tap_step) {
#pragma unroll
            for (int c = 0; c < KCH; ++c, dw += (uint64_t)(WBLK >> 4)) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                if (do_mma) { if (PAIR) tc2_mma_f16(tacc, da_tap + (uint64_t)c * chunk_step + uint64_t(2 * ks), dw + uint64_t(2 * ks), idesc, acc);
                             else tc_mma_f16(tacc, da_tap + (uint64_t)c * chunk_step + uint64_t(2 * ks), dw + uint64_t(2 * ks), idesc, acc); }
                acc = 1u;
              }
            }
          }
        } else {
          for (int tap = 0; tap < k; ++tap, da_tap += tap_step) {
#pragma unroll
            for (int c = 0; c < KCH; ++c) {
              mbar_wait(wr_full(slot), slot_ph);
              tc_fence_after();
              const uint64_t dw = make_smem_desc<BKC>(ring + slot * WBLK);
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                if (do_mma) { if (PAIR) tc2_mma_f16(tacc, da_tap + (uint64_t)c * chunk_step + uint64_t(2 * ks), dw + uint64_t(2 * ks), idesc, acc);
                             else tc_mma_f16(tacc, da_tap + (uint64_t)c * chunk_step + uint64_t(2 * ks), dw + uint64_t(2 * ks), idesc, acc); }
                acc = 1u;
              }
              commit(wr_empty(slot));
              if (++slot == SW) { slot = 0; slot_ph ^= 1u; }
            }
          }
        }
      };
      RingPos x1(NX), d1(ND), x2(NX), d2(ND), t2(NT ? NT : 1);     // stage-1 / stage-2 positions
      auto issue1 = [&](int i) {
        const int xb = x1.idx, db = d1.idx;
        if (mt == 0) RP_TRACE(16, i);
        mbar_wait(x_full(xb), x1.ph);
        if (mt == 0) RP_TRACE(13, i);
        mbar_wait(d1_empty(db), d1.ph ^ 1u);
        tc_fence_after();
        if (mt == 0) RP_TRACE(1, i);
        const uint64_t da = make_smem_desc<BKC>(base + sp.x_off + xb * sp.xb + (uint32_t)(mt * 128) * RB);
        run_conv(d1_col(db, mt), da, (uint64_t)((uint32_t)a.dil * RB >> 4), x_chunk_step, w1_res, a.stream1 != 0, 0u);
        commit(d1_full(db));
        if (NT) commit(x_empty(xb));      // stage 1 was the last tensor-pipe reader of the input tile
        if (mt == 0) RP_TRACE(2, i);
        x1.next(); d1.next();
      };
      auto issue2 = [&](int i) {
        const int xb = x2.idx, db = d2.idx;
        const int tb = NT ? t2.idx : xb;                       // stage-2 operand buffer
        if (mt == 0) RP_TRACE(14, i);
        mbar_wait(t_full(tb), NT ? t2.ph : x2.ph);
        if (mt == 0) RP_TRACE(15, i);
        mbar_wait(d2_init(db), d2.ph);
        tc_fence_after();
        if (mt == 0) RP_TRACE(3, i);
        const uint32_t t_base = NT ? base + sp.t_off + tb * sp.tb : base + sp.x_off + xb * sp.xb;
        const uint64_t da = make_smem_desc<BKC>(t_base + (uint32_t)(mt * 128) * RB);
        run_conv(d2_col(db, mt), da, (uint64_t)(RB >> 4), NT ? t_chunk_step : x_chunk_step, w2_res, a.stream2 != 0, 1u);
        commit(NT ? t_empty(tb) : x_empty(xb));
        commit(d2_full(db));
        if (mt == 0) RP_TRACE(4, i);
        x2.next(); d2.next(); t2.next();
      };
      // pipelined: stage 1 of the next live tile is issued before stage 2 of the current one -- the tensor pipe
      // works on it while epilogue group 1 converts this tile's accumulator into the stage-2 operand
      TilePos tp(first_tile, tile_step, a.tiles_per_item);
      int pending = -1;
      for (int i = 0; i < n_local; ++i, tp.next()) {
        if (skip_tile(tp)) continue;
        issue1(i);
        if (a.pipelined) { if (pending >= 0) issue2(pending); pending = i; }
        else issue2(i);
      }
      if (pending >= 0) issue2(pending);
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ===== stage-1 group: seed D2 with residual + bias2, then epilogue 1 (D1 -> stage-2 operand) =====
    const int q = warp & 3;
    const int mt = (warp - 4) >> 2;      // this warp's M-tile
    const float slope = a.slope, inv_slope = a.inv_slope;
    RingPos xr(NX), dr(ND), tr(NT ? NT : 1);
    TilePos tp(first_tile, tile_step, a.tiles_per_item);
    for (int i = 0; i < n_local; ++i, tp.next()) {
      if (skip_tile(tp)) continue;
      const int b = tp.b, tt = tp.tt;
      const int xb = xr.idx, db = dr.idx;
      const int o0 = tt * a.R_out;
      int len_b = a.L;
      if (b >= a.B) len_b = 0;                                   // the padding tile of an odd pair
      else if (a.lens != nullptr) len_b = min(__ldg(a.lens + b), a.L);
      const uint32_t xa = base + sp.x_off + xb * sp.xb;
      const int tb = NT ? tr.idx : xb;
      // seed: D2 <- residual (inverse LeakyReLU of the input tile's rows) + bias2
      auto seed = [&]() {
        mbar_wait(PAIR ? x_seen(xb) : x_full(xb), xr.ph);
        mbar_wait(d2_empty(db), dr.ph ^ 1u);
        tc_fence_after();
        if (warp == 4 && lane == 0) RP_TRACE(5, i);
        if (!(a.dbg & 16)) {
          const uint32_t arow = (uint32_t)(mt * 128 + q * 32 + lane + a.p1 + a.p2);
          const uint32_t swz = (BKC == 64) ? (arow & 7u) : ((arow >> 1) & 3u);
          const uint32_t rbase = xa + arow * RB;
          const uint32_t tcol = d2_col(db, mt) + (uint32_t(q * 32) << 16);
#pragma unroll
          for (int cc = 0; cc < C / 16; ++cc) {
            const int c = (cc * 2) / UPC, u = (cc * 2) % UPC;
            const uint4 t0 = lds128(rbase + c * chunk_bytes + (((uint32_t)u ^ swz) << 4));
            const uint4 t1 = lds128(rbase + c * chunk_bytes + (((uint32_t)(u + 1) ^ swz) << 4));
            const uint32_t w[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            uint32_t r[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v0, v1;
              unpack2<BF16>(w[e], v0, v1);
              // inverse LeakyReLU without a compare / select: inv_slope >= 1, so v * inv_slope <= v exactly when v <= 0
              v0 = fminf(v0, v0 * inv_slope) + bias_s[C + cc * 16 + 2 * e];
              v1 = fminf(v1, v1 * inv_slope) + bias_s[C + cc * 16 + 2 * e + 1];
              r[2 * e] = __float_as_uint(v0); r[2 * e + 1] = __float_as_uint(v1);
            }
            tc_st16(tcol + (uint32_t)(cc * 16), r);
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive_leader_relaxed(d2_init(db));
          if (NT) mbar_arrive(x_empty(xb));      // this warp's residual rows are in TMEM
        }
        if (warp == 4 && lane == 0) RP_TRACE(6, i);
      };
      // epilogue 1: D1 -> + bias1 -> LeakyReLU -> zero outside the sequence -> 16-bit stage-2 operand
      auto epi1 = [&]() {
        mbar_wait(d1_full(db), dr.ph);
        tc_fence_after();
        if (warp == 4 && lane == 0) RP_TRACE(7, i);
        if (NT) mbar_wait(t_empty(tb), tr.ph ^ 1u);   // stage 2 of tile i - NT has read the buffer
        else group_sync(1);   // every warp of the group has read its residual rows: the tile may be overwritten
        if (warp == 4 && lane == 0) RP_TRACE(8, i);
        const uint32_t t_cb = NT ? t_chunk_bytes : chunk_bytes;
        if (!(a.dbg & 8)) {
          const uint32_t m = (uint32_t)(mt * 128 + q * 32 + lane);
          const int trow = o0 - a.p2 + (int)m;
          const bool valid = trow >= 0 && trow < len_b;
          const bool all_valid = !(a.dbg & 64) && __all_sync(0xffffffffu, valid);   // true for all but the tiles at a sequence's ends
          const uint32_t swz = (BKC == 64) ? (m & 7u) : ((m >> 1) & 3u);
          const uint32_t rbase = (NT ? base + sp.t_off + tb * sp.tb : xa) + m * RB;
          const uint32_t tcol = d1_col(db, mt) + (uint32_t(q * 32) << 16);
#pragma unroll
          for (int c32 = 0; c32 < C / 32; ++c32) {
            // one TMEM round trip per 32 columns (two loads in flight, one wait)
            uint32_t r[2][16];
            tc_ld16(tcol + (uint32_t)(c32 * 32), r[0]);
            tc_ld16(tcol + (uint32_t)(c32 * 32 + 16), r[1]);
            tc_wait_ld();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cc = c32 * 2 + h;
              float v[16];
#pragma unroll
              if (all_valid) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float t = __uint_as_float(r[h][e]) + bias_s[cc * 16 + e];
                  v[e] = fmaxf(t, t * slope);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float t = __uint_as_float(r[h][e]) + bias_s[cc * 16 + e];
                  v[e] = valid ? fmaxf(t, t * slope) : 0.f;
                }
              }
              const int c = (cc * 2) / UPC, u = (cc * 2) % UPC;
              sts128(rbase + c * t_cb + (((uint32_t)u ^ swz) << 4),
                     make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7])));
              sts128(rbase + c * t_cb + (((uint32_t)(u + 1) ^ swz) << 4),
                     make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]), pack2<BF16>(v[14], v[15])));
            }
          }
        }
        if (warp == 4 && lane == 0) RP_TRACE(9, i);
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { arrive_leader_relaxed(d1_empty(db)); arrive_leader(t_full(tb)); }
        if (warp == 4 && lane == 0) RP_TRACE(10, i);
      };
      // With its own operand buffer, epilogue 1 of tile i depends only on stage 1 of tile i, while the seed has to
      // wait for the drain of tile i - 2 (the D2 buffer): doing the epilogue first takes ~1000 cycles of this group's
      // work out of the D2 buffer's seed -> stage 2 -> drain cycle.  In place, the seed must read the rows first.
      // With a single accumulator set (C = 128) the D2 buffer is free long before stage 1 ends: seed first, under
      // stage 1's MMAs.
      if (NT && ND == 2) { epi1(); seed(); } else { seed(); epi1(); }
      xr.next(); dr.next(); tr.next();
    }
  } else if (warp >= 12) {
    // ===== stage-2 group: epilogue 2 (D2 -> other branches, scale, mask, activation -> TMA store) =====
    const int q = warp & 3;
    const int mt = (warp - 12) >> 2;     // this warp's M-tile
    const uint32_t stg0 = base + sp.stg_off + (uint32_t)(warp - 12) * a.NSTG * 32u * RB;
    const uint32_t swz = (BKC == 64) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
    const float scale = a.out_scale, aslope = a.out_slope_eff;
    const bool has_res = a.res2 != nullptr || a.res3 != nullptr;
    RingPos dr(ND), sr(a.NSTG);
    TilePos tp(first_tile, tile_step, a.tiles_per_item);
    for (int i = 0; i < n_local; ++i, tp.next()) {
      const int b = tp.b, tt = tp.tt;
      const int o0 = tt * a.R_out;
      if (skip_tile(tp)) {
        // all padding: zero this warp's 32 rows (16-byte stores, a row's units on adjacent lanes)
        if (b < a.B) {
          constexpr int U = C / 8;                              // 16-byte units per row
          const int r0 = o0 + mt * 128 + q * 32;
          const int nrows = min(32, min(a.R_out - (mt * 128 + q * 32), a.L - r0));
          uint16_t* yb = reinterpret_cast<uint16_t*>(a.y) + ((long long)b * a.L + r0) * a.y_ld;
          for (int idx = lane; idx < nrows * U; idx += 32)
            *reinterpret_cast<uint4*>(yb + (long long)(idx / U) * a.y_ld + (idx % U) * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }
      const int db = dr.idx;
      int len_b = a.L;
      if (b >= a.B) len_b = 0;                                   // the padding tile of an odd pair
      else if (a.lens != nullptr) len_b = min(__ldg(a.lens + b), a.L);
      mbar_wait(d2_full(db), dr.ph);
      tc_fence_after();
      if (warp == 12 && lane == 0) RP_TRACE(11, i);
      {
        const int o = mt * 128 + q * 32 + lane;
        const int grow = o0 + o;
        const bool masked = grow >= len_b;
        // most tiles: nothing masked, no scaling (only the last pair of a stage averages the MRF branches)
        const bool plain = !(a.dbg & 64) && scale == 1.f && !__any_sync(0xffffffffu, masked);
        const int rows_here = min(32, a.R_out - (mt * 128 + q * 32));   // warp-uniform
        const uint32_t tcol = d2_col(db, mt) + (uint32_t(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < KCH; ++c) {
          const uint32_t stg = stg0 + (uint32_t)sr.idx * 32u * RB;
          sr.next();
          if (lane == 0) { if (a.NSTG > 1) bulk_wait_read1(); else bulk_wait_read0(); }
          __syncwarp();
#pragma unroll
          for (int c32 = 0; c32 < BKC / 32; ++c32) {
            // one TMEM round trip per 32 columns; the other MRF branches' rows are fetched while it is in flight
            const int cb = c * BKC + c32 * 32;
            uint32_t r[2][16];
            tc_ld16(tcol + (uint32_t)cb, r[0]);
            tc_ld16(tcol + (uint32_t)(cb + 16), r[1]);
            uint4 g2[4], g3[4];
            const bool add_res = has_res && o < a.R_out && grow < a.L && b < a.B;
            if (add_res) {
              const long long rrow = (long long)b * a.L + grow;
              if (a.res2 != nullptr) {
                const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.res2) + rrow * a.res2_ld + cb);
#pragma unroll
                for (int j = 0; j < 4; ++j) g2[j] = __ldg(p + j);
              }
              if (a.res3 != nullptr) {
                const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.res3) + rrow * a.res3_ld + cb);
#pragma unroll
                for (int j = 0; j < 4; ++j) g3[j] = __ldg(p + j);
              }
            }
            tc_wait_ld();
            if (warp == 12 && lane == 0 && c == 0 && c32 == 0) RP_TRACE(17, i);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cc = c32 * 2 + h;
              float v[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[h][e]);
              if (add_res) {
                if (a.res2 != nullptr) {
                  const uint4 t0 = g2[2 * h], t1 = g2[2 * h + 1];
                  unpack_add<BF16>(t0.x, v[0], v[1]); unpack_add<BF16>(t0.y, v[2], v[3]);
                  unpack_add<BF16>(t0.z, v[4], v[5]); unpack_add<BF16>(t0.w, v[6], v[7]);
                  unpack_add<BF16>(t1.x, v[8], v[9]); unpack_add<BF16>(t1.y, v[10], v[11]);
                  unpack_add<BF16>(t1.z, v[12], v[13]); unpack_add<BF16>(t1.w, v[14], v[15]);
                }
                if (a.res3 != nullptr) {
                  const uint4 t0 = g3[2 * h], t1 = g3[2 * h + 1];
                  unpack_add<BF16>(t0.x, v[0], v[1]); unpack_add<BF16>(t0.y, v[2], v[3]);
                  unpack_add<BF16>(t0.z, v[4], v[5]); unpack_add<BF16>(t0.w, v[6], v[7]);
                  unpack_add<BF16>(t1.x, v[8], v[9]); unpack_add<BF16>(t1.y, v[10], v[11]);
                  unpack_add<BF16>(t1.z, v[12], v[13]); unpack_add<BF16>(t1.w, v[14], v[15]);
                }
              }
              if (plain) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], v[e] * aslope);
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float t = masked ? 0.f : v[e] * scale;
                  v[e] = fmaxf(t, t * aslope);
                }
              }
              const uint32_t o_lo = stg + (uint32_t)lane * RB + ((((uint32_t)(2 * cc)) ^ swz) << 4);
              const uint32_t o_hi = stg + (uint32_t)lane * RB + ((((uint32_t)(2 * cc + 1)) ^ swz) << 4);
              sts128(o_lo, make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7])));
              sts128(o_hi, make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]), pack2<BF16>(v[14], v[15])));
            }
          }
          if (warp == 12 && lane == 0 && c == KCH - 1) RP_TRACE(18, i);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (a.dbg & 2) {}
            else if (rows_here >= 32) tma_store_3d(&maps.y, stg, c * BKC, o0 + mt * 128 + q * 32, b);
            else if (rows_here > 0) tma_store_3d(&maps.y_tail, stg, c * BKC, o0 + mt * 128 + q * 32, b);
            bulk_commit();
          }
          __syncwarp();
        }
      }
      if (warp == 12 && lane == 0) RP_TRACE(19, i);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(db));
      if (warp == 12 && lane == 0) RP_TRACE(12, i);
      dr.next();
    }
    if (lane == 0) bulk_wait_all0();
  }

  tc_fence_before();
  if (PAIR) cluster_sync();     // the peer may still multicast commits into this CTA / read its operands
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int C, bool BF16>
__global__ void __launch_bounds__(RP_THREADS, 1)
resblock_pair_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ PairArgs a) {
  resblock_pair_body<C, BF16, false>(maps, a);
}
template <int C, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RP_THREADS, 1)
resblock_pair2_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ PairArgs a) {
  resblock_pair_body<C, BF16, true>(maps, a);
}

template <int C, bool BF16>
static int launch_pair2(const PairMaps& maps, const PairArgs& a, size_t smem, cudaStream_t st) {
  ASB_SMEM_OPT_IN(227 * 1024, resblock_pair2_kernel<C, BF16>);
  int clusters = num_sms() / 2;
  const int slots = (a.total_tiles + 1) / 2;
  if (clusters > slots) clusters = slots;
  ASB_CUDA(launch_k(resblock_pair2_kernel<C, BF16>, 2 * clusters, RP_THREADS, smem, st, maps, a));
  return AS_OK;
}

template <int C, bool BF16>
static int launch_pair(const PairMaps& maps, const PairArgs& a, size_t smem, cudaStream_t st) {
  if (a.pair) return launch_pair2<C, BF16>(maps, a, smem, st);
  ASB_SMEM_OPT_IN(227 * 1024, resblock_pair_kernel<C, BF16>);
  int grid = num_sms();
  static const int grid_cap = getenv("ASB_PAIR_GRID") ? atoi(getenv("ASB_PAIR_GRID")) : 0;   // experiments: leave SMs to concurrent streams
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  if (grid > a.total_tiles) grid = a.total_tiles;
  ASB_CUDA(launch_k(resblock_pair_kernel<C, BF16>, grid, RP_THREADS, smem, st, maps, a));
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace asb

// development aid, not part of the public header: copies the hand-off trace of the last ASB_PAIR_DBG=32 launch
extern "C" int as_debug_pair_trace(unsigned long long* host, int n) {
  if (n > 20 * 32) n = 20 * 32;
  return asb::check_cuda(cudaMemcpyFromSymbol(host, asb::g_pair_trace, sizeof(unsigned long long) * n), "trace");
}

extern "C" int as_hifigan_resblock_pair(const as_resblock_pair_params* p, void* stream) {
  using namespace asb;
  ASB_REQUIRE(p != nullptr, AS_ERR_SHAPE, "as_hifigan_resblock_pair: null params");
  ASB_REQUIRE(p->dtype == AS_F16 || p->dtype == AS_BF16, AS_ERR_DTYPE, "as_hifigan_resblock_pair: dtype must be f16 or bf16");
  ASB_REQUIRE(p->C == 32 || p->C == 64 || p->C == 128, AS_ERR_SHAPE, "as_hifigan_resblock_pair: C=%d not in {32, 64, 128}", p->C);
  ASB_REQUIRE(p->B > 0 && p->L > 0, AS_ERR_SHAPE, "as_hifigan_resblock_pair: non-positive dimension");
  ASB_REQUIRE(p->k >= 1 && (p->k & 1) == 1 && p->k <= 15 && p->dil >= 1 && (p->k - 1) * p->dil <= 64, AS_ERR_SHAPE,
              "as_hifigan_resblock_pair: k=%d dil=%d unsupported (odd k <= 15, (k-1)*dil <= 64)", p->k, p->dil);
  ASB_REQUIRE(p->x && p->w1 && p->w2 && p->b1 && p->b2 && p->y, AS_ERR_SHAPE, "as_hifigan_resblock_pair: null pointer");
  ASB_REQUIRE(p->slope > 0.f && p->slope <= 1.f, AS_ERR_SHAPE, "as_hifigan_resblock_pair: slope must be in (0, 1]");
  ASB_REQUIRE(p->out_act == AS_ACT_NONE || (p->out_act == AS_ACT_LRELU && p->out_slope <= 1.f), AS_ERR_SHAPE,
              "as_hifigan_resblock_pair: out_act must be none or LeakyReLU with slope <= 1");
  auto aligned = [](const void* ptr, long long ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0; };
  ASB_REQUIRE(aligned(p->x, p->x_ld) && aligned(p->y, p->y_ld) && p->x_ld >= p->C && p->y_ld >= p->C, AS_ERR_ALIGN,
              "as_hifigan_resblock_pair: x / y must be 16-byte aligned with row strides that are multiples of 8");
  ASB_REQUIRE((reinterpret_cast<uintptr_t>(p->w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w2) & 15) == 0, AS_ERR_ALIGN,
              "as_hifigan_resblock_pair: weights must be 16-byte aligned");
  ASB_REQUIRE((p->res2 == nullptr || aligned(p->res2, p->res2_ld)) && (p->res3 == nullptr || aligned(p->res3, p->res3_ld)),
              AS_ERR_ALIGN, "as_hifigan_resblock_pair: res2 / res3 alignment");
  int rc = check_arch();
  if (rc != AS_OK) return rc;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return AS_ERR_CUDA;

  const int C = p->C, k = p->k;
  const int BKC = C >= 64 ? 64 : 32, KCH = C / BKC;
  const size_t RB = BKC * 2, WBLK = (size_t)C * RB;
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.B = p->B; a.L = p->L; a.k = k; a.dil = p->dil;
  a.p1 = (k - 1) * p->dil / 2; a.p2 = (k - 1) / 2;
  a.R_out = 256 - (k - 1);
  a.HA = (256 + (k - 1) * p->dil + 15) / 16 * 16;
  a.HB = a.HA / 2;
  a.tail_rows = a.R_out - 224;
  a.tiles_per_item = (p->L + a.R_out - 1) / a.R_out;
  a.total_tiles = p->B * a.tiles_per_item;
  a.ND = C <= 64 ? 2 : 1;
  // shared-memory plan.  C <= 64: the stage-2 operand gets its own buffers (NT = 2, else 1), so an input buffer is
  // held only from its TMA load to the end of stage 1 and the load of tile i + NX overlaps everything after that
  // (measured: with the operand written in place the buffer lived for a whole tile, and the tensor pipe idled for
  // the 2-4k cycles of every load).  Weights are resident when they fit next to that, otherwise conv1's (then both)
  // stream from L2 through an mbarrier ring.  C = 128: one input buffer, operand in place, weights streamed.
  // Remaining room goes to the weight ring, a second staging buffer, then more input buffers (up to 8).
  const size_t cap = 227 * 1024 - 1024;
  const size_t nblk = (size_t)k * KCH;
  int best_found = 0;
  const int mode_lo = env_int("ASB_PAIR_MODE", 0);
  const int nt_force = env_int("ASB_PAIR_NT", -1);
  // CTA pairs (cta_group::2): each CTA holds half of every weight block.  Needs at least two tiles and the
  // separate-operand-buffer mode (the peer CTA's epilogue group cannot wait on the leader's x_full barrier).
  // Measured (profiles/r02_pair_2cta_ab.txt, same-run A/B): the MMA phases get 34 % faster (59 vs 89 cycles per
  // instruction at N = 128), but every cross-CTA hand-off that publishes memory costs ~1300 cycles (cluster-scope
  // release on the remote arrive; TMEM-only hand-offs arrive relaxed), so the pair pays off only where a tile carries
  // enough MMA work: C = 64: k = 7 -19 %, k = 11 -5 %; C = 128: k = 7 -2 %, k = 11 -4 %; slower for k = 3 and C = 32.
  // ASB_PAIR_2CTA = 0 / 1 forces it off / on (where a plan fits).
  const int pair_env = env_int("ASB_PAIR_2CTA", -1);
  a.pair = (pair_env >= 0 ? pair_env != 0 : (C >= 64 && k >= 7)) && a.total_tiles >= 2 ? 1 : 0;
  auto fits = [&](int nx, int nt, int nstg, int sw, int s1, int s2) {
    return pair_smem(C, a.HA, k, nx, nt, nstg, sw, s1, s2, a.pair).total <= cap;
  };
  for (int nt = ((C <= 64 || a.pair) ? 2 : 0); nt >= (a.pair ? 1 : 0) && !best_found; --nt) {
    if (nt_force >= 0 && nt != nt_force) continue;
    for (int mode = mode_lo; mode < 3 && !best_found; ++mode) {   // 0: resident, 1: conv1 streamed, 2: both streamed
      a.stream1 = mode >= 1; a.stream2 = mode >= 2;
      for (int nx = 2; nx >= 1 && !best_found; --nx) {
        if (nx == 1 && (mode == 0 || (nt > 0 && C <= 64))) continue;   // prefer streaming / fewer operand buffers over one input buffer
                                                                       // (C = 128 in pair mode only fits one input + one operand buffer)
        a.NX = nx; a.NT = nt; a.NSTG = 1;
        a.SW = mode == 0 ? 0 : 3;
        if (!fits(a.NX, a.NT, a.NSTG, a.SW, a.stream1, a.stream2)) continue;
        best_found = 1;
        if (mode > 0) {
          const int want = (int)((a.stream1 ? nblk : 0) + (a.stream2 ? nblk : 0));
          const int sw_max = nt > 0 ? 8 : 16;
          while (a.SW < sw_max && a.SW < want && fits(a.NX, a.NT, a.NSTG, a.SW + 1, a.stream1, a.stream2)) ++a.SW;
        }
        if (fits(a.NX, a.NT, 2, a.SW, a.stream1, a.stream2)) a.NSTG = 2;
        const int nx_max = nt > 0 ? 6 : 3;
        while (a.NX < nx_max && (mode <= 1 || nt > 0) && fits(a.NX + 1, a.NT, a.NSTG, a.SW, a.stream1, a.stream2)) ++a.NX;
      }
    }
  }
  if (!best_found && a.pair) {      // no separate-buffer plan fits: single-CTA kernel with the operand in place
    a.pair = 0;
    for (int mode = 1; mode < 3 && !best_found; ++mode)
      for (int nx = 2; nx >= 1 && !best_found; --nx) {
        a.stream1 = 1; a.stream2 = mode >= 2; a.NX = nx; a.NT = 0; a.NSTG = 1; a.SW = 3;
        if (fits(a.NX, 0, 1, a.SW, a.stream1, a.stream2)) best_found = 1;
      }
  }
  if (!best_found) {
    a.stream1 = a.stream2 = 1; a.NX = 1; a.NT = 0; a.NSTG = 1; a.SW = 2;
    ASB_REQUIRE(fits(1, 0, 1, 2, 1, 1), AS_ERR_SHAPE, "as_hifigan_resblock_pair: tile does not fit in shared memory");
  }
  // tuning overrides (tools/prof_pair.py)
  a.NX = env_int("ASB_PAIR_NX", a.NX); a.NSTG = env_int("ASB_PAIR_NSTG", a.NSTG); a.SW = env_int("ASB_PAIR_SW", a.SW);
  // pipelined issue order (stage 1 of tile i + 1 before stage 2 of tile i) needs two accumulator sets and two tiles
  // in flight; the weight ring is filled in the same order
  a.pipelined = (a.ND == 2 && a.NX >= 2) ? 1 : 0;
  a.pipelined = env_int("ASB_PAIR_PIPE", a.pipelined);
  if (a.ND < 2 || a.NX < 2) a.pipelined = 0;
  const PairSmem sp = pair_smem(C, a.HA, k, a.NX, a.NT, a.NSTG, a.SW, a.stream1, a.stream2, a.pair);
  ASB_REQUIRE(!(a.pair && a.NT == 0), AS_ERR_SHAPE, "as_hifigan_resblock_pair: CTA pairs need separate operand buffers");
  ASB_REQUIRE(sp.total <= cap && a.NX >= 1 && a.NX <= 8 && a.NT >= 0 && a.NT <= 2 && a.NSTG >= 1 && a.NSTG <= 2 && a.SW <= 16 &&
                  (!(a.stream1 || a.stream2) || a.SW >= 2),
              AS_ERR_SHAPE, "as_hifigan_resblock_pair: invalid shared-memory plan (%u bytes)", sp.total);
  if (env_int("ASB_PAIR_VERBOSE", 0))
    fprintf(stderr, "pair C=%d k=%d dil=%d: 2cta=%d NX=%d NT=%d ND=%d NSTG=%d SW=%d stream=%d%d pipelined=%d smem=%u\n", C, k, p->dil,
            a.pair, a.NX, a.NT, a.ND, a.NSTG, a.SW, a.stream1, a.stream2, a.pipelined, sp.total);

  const uint32_t fmt = (p->dtype == AS_BF16) ? 1u : 0u;
  a.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(C >> 3) << 17) | (uint32_t((a.pair ? 256 : 128) >> 4) << 24);
  a.b1 = p->b1; a.b2 = p->b2;
  a.res2 = p->res2; a.res2_ld = p->res2_ld; a.res3 = p->res3; a.res3_ld = p->res3_ld;
  a.slope = p->slope; a.inv_slope = 1.0f / p->slope;
  a.out_scale = p->out_scale;
  a.out_slope_eff = p->out_act == AS_ACT_LRELU ? p->out_slope : 1.0f;
  a.lens = p->lens;
  a.y = p->y; a.y_ld = p->y_ld;
  a.dbg = env_int("ASB_PAIR_DBG", 0);

  const CUtensorMapDataType dt = p->dtype == AS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = BKC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  PairMaps maps;
  memset(&maps, 0, sizeof(maps));
  auto mk3 = [&](CUtensorMap* m, const void* ptr, long long ld, int rows) -> bool {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)p->L, (cuuint64_t)p->B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * p->L};
    cuuint32_t box[3] = {(cuuint32_t)BKC, (cuuint32_t)rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto mkw = [&](CUtensorMap* m, const void* ptr) -> bool {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)k * C};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)BKC, (cuuint32_t)(a.pair ? C / 2 : C)};   // pair mode: each CTA loads its half of a block
    cuuint32_t es[2] = {1, 1};
    return enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  bool ok = mk3(&maps.x, p->x, p->x_ld, a.HB) && mkw(&maps.w1, p->w1) && mkw(&maps.w2, p->w2) &&
            mk3(&maps.y, p->y, p->y_ld, 32) && mk3(&maps.y_tail, p->y, p->y_ld, a.tail_rows);
  ASB_REQUIRE(ok, AS_ERR_CUDA, "as_hifigan_resblock_pair: cuTensorMapEncodeTiled failed");

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = sp.total + 1024;
  const bool bf = p->dtype == AS_BF16;
  if (C == 32) return bf ? launch_pair<32, true>(maps, a, smem, st) : launch_pair<32, false>(maps, a, smem, st);
  if (C == 64) return bf ? launch_pair<64, true>(maps, a, smem, st) : launch_pair<64, false>(maps, a, smem, st);
  return bf ? launch_pair<128, true>(maps, a, smem, st) : launch_pair<128, false>(maps, a, smem, st);
}
