// Memory-bound kernels of the synthesis path: embedding, LayerNorm, InstanceNorm statistics,
// AdaIN(+LeakyReLU, + depthwise ConvTranspose "pool"), nearest upsample, length regulation,
// small-Cin direct convolution, depthwise convolution, pooling, layout changes.
// All of them are coalesced along the channel axis of the channels-last activations and
// vectorised where the row pitch allows it; reductions use warp shuffles.
#include "common.cuh"

namespace asb {

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// embedding * scale  (Utils/RelTransformerEnc.py:372-373)
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table,
                             int n_vocab, int B, int T, int C, float scale,
                             const int* __restrict__ lens, float* out32, void* out16, int dt16) {
  const long long row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row % T;
  const bool valid = lens == nullptr || t < lens[b];
  long long id = tok[row];
  if (id < 0 || id >= n_vocab) id = 0;
  const float* src = table + id * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = valid ? src[c] * scale : 0.f;
    if (out32) out32[row * C + c] = v;
    if (out16) reinterpret_cast<uint16_t*>(out16)[row * C + c] = to16(v, dt16);
  }
}

// ---------------------------------------------------------------------------------------------
// channel LayerNorm, one warp per row (C <= 1024)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAX_PER_LANE = 32;

__global__ void layernorm_kernel(const void* __restrict__ x, int xdt, long long x_ld, long long rows,
                                 int T, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int act, float slope,
                                 const int* __restrict__ lens, void* oa, int oadt, long long oa_ld,
                                 void* ob, int obdt, long long ob_ld) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row % T;
  const bool valid = lens == nullptr || t < lens[b];
  float v[LN_MAX_PER_LANE];
  float s = 0.f;
  const int n = (C + 31) / 32;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    v[i] = 0.f;
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) { v[i] = ldany(x, row * x_ld + c, xdt); s += v[i]; }
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) { float d = v[i] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      int c = lane + 32 * i;
      if (c < C) {
        float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
        y = valid ? apply_act(y, act, slope) : 0.f;
        if (oa) stany(oa, row * oa_ld + c, y, oadt);
        if (ob) stany(ob, row * ob_ld + c, y, obdt);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm statistics: block = (b, 32 channels), 32 x 8 threads, two passes (mean, then
// centred second moment) like the reference's biased variance.
// ---------------------------------------------------------------------------------------------
__global__ void instnorm_stats_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T,
                                      int C, const int* __restrict__ lens, float eps,
                                      float* __restrict__ stats) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const int len = lens ? min(lens[b], T) : T;
  const long long base = (long long)b * T * x_ld;
  float s = 0.f;
  if (c < C)
    for (int t = ty; t < len; t += 8) s += ldany(x, base + (long long)t * x_ld + c, xdt);
  red[ty][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
  const float mean = len > 0 ? tot / len : 0.f;
  __syncthreads();
  float q = 0.f;
  if (c < C)
    for (int t = ty; t < len; t += 8) {
      float d = ldany(x, base + (long long)t * x_ld + c, xdt) - mean;
      q += d * d;
    }
  red[ty][threadIdx.x] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += red[i][threadIdx.x];
    var = len > 0 ? var / len : 0.f;
    stats[((long long)b * C + c) * 2 + 0] = mean;
    stats[((long long)b * C + c) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------------------------------------
// AdaIN + LeakyReLU (+ depthwise ConvTranspose1d k3 s2 p1 op1)
// ---------------------------------------------------------------------------------------------
__global__ void adain_apply_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                   int C, const float* __restrict__ stats,
                                   const float* __restrict__ gb, long long gb_ld, float slope,
                                   const int* __restrict__ lens, const float* __restrict__ up_w,
                                   const float* __restrict__ up_b, void* out, int odt,
                                   long long out_ld) {
  const int To = up_w ? 2 * T : T;
  const long long total = (long long)B * To * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long r = i / C;
    const int to = r % To;
    const int b = r / To;
    const int len = lens ? min(lens[b], T) : T;
    const float mean = stats[((long long)b * C + c) * 2], rstd = stats[((long long)b * C + c) * 2 + 1];
    const float g = 1.f + gb[(long long)b * gb_ld + c], be = gb[(long long)b * gb_ld + C + c];
    auto a_at = [&](int t) -> float {
      if (t >= len) return 0.f;
      float v = ldany(x, ((long long)b * T + t) * x_ld + c, xdt);
      v = (v - mean) * rstd * g + be;
      return v > 0.f ? v : v * slope;
    };
    float y;
    if (up_w == nullptr) {
      y = a_at(to);
    } else {
      const int m = to >> 1;
      if (m >= len) y = 0.f;
      else if ((to & 1) == 0) y = a_at(m) * up_w[c * 3 + 1] + up_b[c];
      else y = a_at(m) * up_w[c * 3 + 2] + a_at(m + 1) * up_w[c * 3 + 0] + up_b[c];
    }
    stany(out, ((long long)b * To + to) * out_ld + c, y, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// nearest upsample along T
// ---------------------------------------------------------------------------------------------
__global__ void repeat_rows_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                   int C, int rep, const int* __restrict__ lens, void* out, int odt,
                                   long long out_ld) {
  const int To = T * rep;
  const long long total = (long long)B * To * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long r = i / C;
    const int to = r % To;
    const int b = r / To;
    const int t = to / rep;
    const int len = lens ? min(lens[b], T) : T;
    float v = t < len ? ldany(x, ((long long)b * T + t) * x_ld + c, xdt) : 0.f;
    stany(out, ((long long)b * To + to) * out_ld + c, v, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// length regulation: one block per item builds frame->token map by a prefix sum, then gathers
// ---------------------------------------------------------------------------------------------
__global__ void length_regulate_kernel(const void* __restrict__ x, int xdt, long long x_ld, int Tt,
                                       int C, const int* __restrict__ dur,
                                       const int* __restrict__ lens_t, int rep, int To, void* out,
                                       int odt, long long out_ld, int* __restrict__ out_lens) {
  extern __shared__ int cum[];  // [Tt + 1] inclusive prefix sums, cum[0] = 0
  const int b = blockIdx.x;
  const int nt = lens_t ? min(lens_t[b], Tt) : Tt;
  if (threadIdx.x == 0) {
    int s = 0;
    cum[0] = 0;
    for (int j = 0; j < nt; ++j) { s += max(dur[(long long)b * Tt + j], 0); cum[j + 1] = s; }
  }
  __syncthreads();
  const int L = cum[nt];
  if (threadIdx.x == 0 && out_lens) out_lens[b] = min(L * rep, To);
  // each warp walks output rows; binary search of the token for the row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < To; r += nw) {
    const int fr = r / rep;
    int tok = -1;
    if (fr < L) {
      int lo = 0, hi = nt;  // find largest j with cum[j] <= fr
      while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cum[mid] <= fr) lo = mid; else hi = mid; }
      tok = lo;
    }
    for (int c = lane; c < C; c += 32) {
      float v = tok >= 0 ? ldany(x, ((long long)b * Tt + tok) * x_ld + c, xdt) : 0.f;
      stany(out, ((long long)b * To + r) * out_ld + c, v, odt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// direct convolution, tiny Cin
// ---------------------------------------------------------------------------------------------
constexpr int CS_MAX_TAPS = 32;
struct SmallTaps { int dt[CS_MAX_TAPS]; int df[CS_MAX_TAPS]; };

__global__ void conv_small_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T,
                                  int F, int Cin, const float* __restrict__ w,
                                  const float* __restrict__ bias, int ntaps, SmallTaps taps,
                                  int Cout, const int* __restrict__ lens, void* yr, int yrdt,
                                  long long yr_ld, void* ya, int yadt, long long ya_ld, int act,
                                  float slope) {
  const long long total = (long long)B * T * F * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const long long row = i / Cout;
    const int f = row % F;
    const int t = (row / F) % T;
    const int b = row / ((long long)F * T);
    float acc = bias ? bias[co] : 0.f;
    for (int j = 0; j < ntaps; ++j) {
      const int ti = t + taps.dt[j], fi = f + taps.df[j];
      if (ti < 0 || ti >= T || fi < 0 || fi >= F) continue;
      const long long xr = (((long long)b * T + ti) * F + fi) * x_ld;
      const float* wj = w + ((long long)j * Cout + co) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc += ldany(x, xr + ci, xdt) * wj[ci];
    }
    if (lens && t >= lens[b]) acc = 0.f;
    if (yr) stany(yr, row * yr_ld + co, acc, yrdt);
    if (ya) stany(ya, row * ya_ld + co, apply_act(acc, act, slope), yadt);
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise convolution with stride (optionally GLU on a 2C-channel input)
// ---------------------------------------------------------------------------------------------
__global__ void dwconv_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F,
                              int C, int glu, const float* __restrict__ w,
                              const float* __restrict__ bias, int kt, int kf, int st, int sf, int pt,
                              int pf, int To, int Fo, const int* __restrict__ lens_in,
                              const int* __restrict__ lens_out, int act, float slope, void* out,
                              int odt, long long out_ld) {
  const long long total = (long long)B * To * Fo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long row = i / C;
    const int fo = row % Fo;
    const int to = (row / Fo) % To;
    const int b = row / ((long long)Fo * To);
    const int len_in = lens_in ? min(lens_in[b], T) : T;
    float acc = bias ? bias[c] : 0.f;
    for (int jt = 0; jt < kt; ++jt) {
      const int ti = to * st + jt - pt;
      if (ti < 0 || ti >= len_in) continue;
      for (int jf = 0; jf < kf; ++jf) {
        const int fi = fo * sf + jf - pf;
        if (fi < 0 || fi >= F) continue;
        const long long xr = (((long long)b * T + ti) * F + fi) * x_ld;
        float v = ldany(x, xr + c, xdt);
        if (glu) { float g = ldany(x, xr + C + c, xdt); v = v / (1.f + __expf(-g)); }
        acc += v * w[(jt * kf + jf) * C + c];
      }
    }
    float y = apply_act(acc, act, slope);
    if (lens_out && to >= lens_out[b]) y = 0.f;
    stany(out, row * out_ld + c, y, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// average pooling with replicate padding of the last T position
// ---------------------------------------------------------------------------------------------
__global__ void avgpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B, int T, int F,
                               int C, int pt, int pf, int To, int Fo, void* out, int odt,
                               long long out_ld) {
  const long long total = (long long)B * To * Fo * C;
  const float inv = 1.f / (pt * pf);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long row = i / C;
    const int fo = row % Fo;
    const int to = (row / Fo) % To;
    const int b = row / ((long long)Fo * To);
    float acc = 0.f;
    for (int jt = 0; jt < pt; ++jt) {
      const int ti = min(to * pt + jt, T - 1);  // replicate the last column when T is odd
      for (int jf = 0; jf < pf; ++jf) {
        const int fi = fo * pf + jf;
        acc += ldany(x, (((long long)b * T + ti) * F + fi) * x_ld + c, xdt);
      }
    }
    stany(out, row * out_ld + c, acc * inv, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(eval) affine -> LeakyReLU -> MaxPool along F
// ---------------------------------------------------------------------------------------------
__global__ void affine_act_maxpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int B,
                                          int T, int F, int C, const float* __restrict__ scale,
                                          const float* __restrict__ shift, float slope, int pf, int Fo,
                                          void* out, int odt, long long out_ld) {
  const long long total = (long long)B * T * Fo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long row = i / C;
    const int fo = row % Fo;
    const long long bt = row / Fo;
    const float sc = scale[c], sh = shift[c];
    float m = -INFINITY;
    for (int j = 0; j < pf; ++j) {
      float v = ldany(x, (bt * F + fo * pf + j) * x_ld + c, xdt) * sc + sh;
      v = v > 0.f ? v : v * slope;
      m = fmaxf(m, v);
    }
    stany(out, row * out_ld + c, m, odt);
  }
}

// ---------------------------------------------------------------------------------------------
// LeakyReLU + global average pool over (T with stride, F)
// ---------------------------------------------------------------------------------------------
__global__ void global_avgpool_kernel(const void* __restrict__ x, int xdt, long long x_ld, int T, int F,
                                      int C, int ts, float slope, void* out, int odt,
                                      long long out_ld) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  int n = 0;
  for (int t = 0; t < T; t += ts)
    for (int f = 0; f < F; ++f) {
      float v = ldany(x, (((long long)b * T + t) * F + f) * x_ld + c, xdt);
      acc += v > 0.f ? v : v * slope;
      ++n;
    }
  stany(out, (long long)b * out_ld + c, acc / n, odt);
}

// ---------------------------------------------------------------------------------------------
// one LSTM step from zero state (EMA_Predictor quirk)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void lstm_onestep_kernel(const float* __restrict__ xp, long long xp_ld, long long rows, int H,
                                    void* out, int odt, long long out_ld) {
  const long long total = rows * 2 * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = i % H;
    const int dir = (i / H) % 2;
    const long long row = i / (2 * H);
    const float* g = xp + row * xp_ld + (long long)dir * 4 * H;
    const float ig = sigmoidf_(g[j]), gg = tanhf(g[2 * H + j]), og = sigmoidf_(g[3 * H + j]);
    const float c = ig * gg;
    stany(out, row * out_ld + dir * H + j, og * tanhf(c), odt);
  }
}

// ---------------------------------------------------------------------------------------------
// log-norm energy: mel fp32 [B, n_mels, T] -> [B, T]
// ---------------------------------------------------------------------------------------------
__global__ void log_norm_kernel(const float* __restrict__ mel, int B, int M, int T, float* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int t = i % T, b = i / T;
  float s = 0.f;
  for (int m = 0; m < M; ++m) {
    float e = expf(mel[((long long)b * M + m) * T + t] * 4.f - 4.f);
    s += e * e;
  }
  out[i] = logf(sqrtf(s));
}

// ---------------------------------------------------------------------------------------------
// channels-first <-> channels-last with cast / affine / masking (tiled transpose)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_cast_kernel(const void* __restrict__ src, int sdt, void* dst, int ddt, int C,
                                      int T, long long cl_ld, int to_cl, const float* __restrict__ sub,
                                      const float* __restrict__ mul, const int* __restrict__ lens) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int len = lens ? min(lens[b], T) : T;
  if (to_cl) {
    // read [c][t] coalesced along t, write [t][c] coalesced along c
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, t = t0 + threadIdx.x;
      float v = 0.f;
      if (c < C && t < T) {
        v = ldany(src, ((long long)b * C + c) * T + t, sdt);
        if (sub) v = (v - sub[c]) * mul[c];
        if (t >= len) v = 0.f;
      }
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int t = t0 + i, c = c0 + threadIdx.x;
      if (c < C && t < T) stany(dst, ((long long)b * T + t) * cl_ld + c, tile[threadIdx.x][i], ddt);
    }
  } else {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int t = t0 + i, c = c0 + threadIdx.x;
      float v = 0.f;
      if (c < C && t < T) {
        v = ldany(src, ((long long)b * T + t) * cl_ld + c, sdt);
        if (sub) v = (v - sub[c]) * mul[c];
        if (t >= len) v = 0.f;
      }
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, t = t0 + threadIdx.x;
      if (c < C && t < T) stany(dst, ((long long)b * C + c) * T + t, tile[threadIdx.x][i], ddt);
    }
  }
}

static inline unsigned ew_grid(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;  // a few waves of the 148 SMs, grid-stride beyond that
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

static inline bool dt_ok(int d) { return d == AS_F16 || d == AS_BF16 || d == AS_F32; }

}  // namespace asb

using namespace asb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int as_embed(const int64_t* tokens, const float* table, int32_t n_vocab, int32_t B,
                        int32_t T, int32_t C, float scale, const int32_t* lens, float* out32,
                        void* out16, int32_t out16_dtype, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(tokens && table && (out32 || out16), AS_ERR_SHAPE, "as_embed: null pointer");
  ASB_REQUIRE(!out16 || out16_dtype == AS_F16 || out16_dtype == AS_BF16, AS_ERR_DTYPE, "as_embed: out16 dtype");
  embed_kernel<<<B * T, 128, 0, ST(stream)>>>(tokens, table, n_vocab, B, T, C, scale, lens, out32, out16, out16_dtype);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_layernorm(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                            int32_t C, const float* gamma, const float* beta, float eps,
                            int32_t act, float slope, const int32_t* lens, void* out_a,
                            int32_t out_a_dtype, int64_t out_a_ld, void* out_b,
                            int32_t out_b_dtype, int64_t out_b_ld, void* stream) {
  const long long rows = (long long)B * T;
  if (rows == 0) return AS_OK;
  ASB_REQUIRE(x && gamma && beta && (out_a || out_b), AS_ERR_SHAPE, "as_layernorm: null pointer");
  ASB_REQUIRE(C > 0 && C <= 32 * LN_MAX_PER_LANE, AS_ERR_SHAPE, "as_layernorm: C=%d unsupported", C);
  ASB_REQUIRE(dt_ok(x_dtype), AS_ERR_DTYPE, "as_layernorm: dtype");
  layernorm_kernel<<<cdiv(rows, 8), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, rows, T, C, gamma, beta, eps, act,
                                                         slope, lens, out_a, out_a_dtype, out_a_ld, out_b,
                                                         out_b_dtype, out_b_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_instnorm_stats(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                 int32_t C, const int32_t* lens, float eps, float* stats,
                                 void* stream) {
  if (B * C == 0) return AS_OK;
  ASB_REQUIRE(x && stats && dt_ok(x_dtype), AS_ERR_SHAPE, "as_instnorm_stats: bad argument");
  dim3 grid(cdiv(C, 32), B), block(32, 8);
  instnorm_stats_kernel<<<grid, block, 0, ST(stream)>>>(x, x_dtype, x_ld, T, C, lens, eps, stats);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_adain_apply(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                              int32_t C, const float* stats, const float* gb, int64_t gb_ld,
                              float slope, const int32_t* lens, const float* up_w,
                              const float* up_b, void* out, int32_t out_dtype, int64_t out_ld,
                              void* stream) {
  const long long total = (long long)B * T * C * (up_w ? 2 : 1);
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && stats && gb && out, AS_ERR_SHAPE, "as_adain_apply: null pointer");
  ASB_REQUIRE(!up_w || up_b, AS_ERR_SHAPE, "as_adain_apply: up_w without up_b");
  adain_apply_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, C, stats, gb, gb_ld, slope,
                                                            lens, up_w, up_b, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_repeat_rows(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                              int32_t C, int32_t rep, const int32_t* lens, void* out,
                              int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = (long long)B * T * rep * C;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && out && rep >= 1, AS_ERR_SHAPE, "as_repeat_rows: bad argument");
  repeat_rows_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, C, rep, lens, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_length_regulate(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t Tt,
                                  int32_t C, const int32_t* dur, const int32_t* lens_t, int32_t rep,
                                  int32_t To, void* out, int32_t out_dtype, int64_t out_ld,
                                  int32_t* out_lens, void* stream) {
  if (B == 0) return AS_OK;
  ASB_REQUIRE(x && dur && out && rep >= 1 && Tt > 0 && To >= 0, AS_ERR_SHAPE, "as_length_regulate: bad argument");
  const size_t smem = (size_t)(Tt + 1) * sizeof(int);
  ASB_REQUIRE(smem <= 48 * 1024, AS_ERR_SHAPE, "as_length_regulate: Tt=%d too large", Tt);
  length_regulate_kernel<<<B, 256, smem, ST(stream)>>>(x, x_dtype, x_ld, Tt, C, dur, lens_t, rep, To, out,
                                                      out_dtype, out_ld, out_lens);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_conv_small(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                             int32_t F, int32_t Cin, const float* w, const float* bias,
                             int32_t ntaps, const int32_t* tap_dt, const int32_t* tap_df,
                             int32_t Cout, const int32_t* lens, void* y_raw, int32_t y_raw_dtype,
                             int64_t y_raw_ld, void* y_act, int32_t y_act_dtype, int64_t y_act_ld,
                             int32_t act, float slope, void* stream) {
  const long long total = (long long)B * T * F * Cout;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && w && (y_raw || y_act) && tap_dt && tap_df, AS_ERR_SHAPE, "as_conv_small: null pointer");
  ASB_REQUIRE(ntaps >= 1 && ntaps <= CS_MAX_TAPS && Cin >= 1 && Cin <= 16, AS_ERR_SHAPE,
              "as_conv_small: ntaps=%d Cin=%d unsupported", ntaps, Cin);
  SmallTaps taps;
  for (int j = 0; j < CS_MAX_TAPS; ++j) { taps.dt[j] = j < ntaps ? tap_dt[j] : 0; taps.df[j] = j < ntaps ? tap_df[j] : 0; }
  conv_small_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, F, Cin, w, bias, ntaps, taps,
                                                           Cout, lens, y_raw, y_raw_dtype, y_raw_ld, y_act,
                                                           y_act_dtype, y_act_ld, act, slope);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_dwconv(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T, int32_t F,
                         int32_t C, int32_t glu, const float* w, const float* bias, int32_t kt,
                         int32_t kf, int32_t st, int32_t sf, int32_t pt, int32_t pf, int32_t To,
                         int32_t Fo, const int32_t* lens_in, const int32_t* lens_out, int32_t act,
                         float slope, void* out, int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = (long long)B * To * Fo * C;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(x && w && out && kt >= 1 && kf >= 1 && st >= 1 && sf >= 1, AS_ERR_SHAPE, "as_dwconv: bad argument");
  dwconv_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, F, C, glu, w, bias, kt, kf, st, sf,
                                                       pt, pf, To, Fo, lens_in, lens_out, act, slope, out,
                                                       out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_avgpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T, int32_t F,
                          int32_t C, int32_t pt, int32_t pf, void* out, int32_t out_dtype,
                          int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && pt >= 1 && pf >= 1, AS_ERR_SHAPE, "as_avgpool: bad argument");
  const int To = (T + pt - 1) / pt, Fo = F / pf;
  const long long total = (long long)B * To * Fo * C;
  if (total == 0) return AS_OK;
  avgpool_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, F, C, pt, pf, To, Fo, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_affine_act_maxpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                     int32_t F, int32_t C, const float* scale, const float* shift,
                                     float slope, int32_t pf, void* out, int32_t out_dtype,
                                     int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && scale && shift && pf >= 1, AS_ERR_SHAPE, "as_affine_act_maxpool: bad argument");
  const int Fo = F / pf;
  const long long total = (long long)B * T * Fo * C;
  if (total == 0) return AS_OK;
  affine_act_maxpool_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(x, x_dtype, x_ld, B, T, F, C, scale, shift,
                                                                   slope, pf, Fo, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_global_avgpool(const void* x, int32_t x_dtype, int64_t x_ld, int32_t B, int32_t T,
                                 int32_t F, int32_t C, int32_t t_stride, float slope, void* out,
                                 int32_t out_dtype, int64_t out_ld, void* stream) {
  ASB_REQUIRE(x && out && t_stride >= 1 && T >= 1 && F >= 1, AS_ERR_SHAPE, "as_global_avgpool: bad argument");
  if (B * C == 0) return AS_OK;
  dim3 grid(cdiv(C, 128), B);
  global_avgpool_kernel<<<grid, 128, 0, ST(stream)>>>(x, x_dtype, x_ld, T, F, C, t_stride, slope, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_lstm_onestep(const float* xproj, int64_t xproj_ld, int64_t rows, int32_t H, void* out,
                               int32_t out_dtype, int64_t out_ld, void* stream) {
  const long long total = rows * 2 * H;
  if (total == 0) return AS_OK;
  ASB_REQUIRE(xproj && out, AS_ERR_SHAPE, "as_lstm_onestep: null pointer");
  lstm_onestep_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(xproj, xproj_ld, rows, H, out, out_dtype, out_ld);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_log_norm(const float* mel, int32_t B, int32_t n_mels, int32_t T, float* out, void* stream) {
  if (B * T == 0) return AS_OK;
  ASB_REQUIRE(mel && out, AS_ERR_SHAPE, "as_log_norm: null pointer");
  log_norm_kernel<<<cdiv((long long)B * T, 128), 128, 0, ST(stream)>>>(mel, B, n_mels, T, out);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}

extern "C" int as_transpose_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype,
                                 int32_t B, int32_t C, int32_t T, int64_t cl_ld,
                                 int32_t to_channels_last, const float* sub, const float* mul,
                                 const int32_t* lens, void* stream) {
  if (B * C * T == 0) return AS_OK;
  ASB_REQUIRE(src && dst && dt_ok(src_dtype) && dt_ok(dst_dtype), AS_ERR_SHAPE, "as_transpose_cast: bad argument");
  ASB_REQUIRE((sub == nullptr) == (mul == nullptr), AS_ERR_SHAPE, "as_transpose_cast: sub/mul must come together");
  dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
  transpose_cast_kernel<<<grid, block, 0, ST(stream)>>>(src, src_dtype, dst, dst_dtype, C, T, cl_ld,
                                                       to_channels_last, sub, mul, lens);
  ASB_CUDA(cudaGetLastError());
  return AS_OK;
}
