/*
 * artspeech_b200 — C ABI of the sm_100a (B200) kernels behind ArtSpeech's batched synthesis
 * forward pass and MAS.  The reference (Zhongxu-Wang/ArtSpeech) is pure Python/PyTorch and has no
 * FFI of its own; each entry point below names the reference code whose
 * arithmetic it replaces.  The host side (artspeech_b200/*.py) binds these with ctypes and keeps
 * the reference's nn.Module constructors / state_dict / forward() signatures.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller unless the
 *     comment says "host".  The library never allocates or frees device memory, never
 *     synchronises the stream, and keeps no global mutable state besides the last-error string.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return 0 (AS_OK) or a negative as_status; as_last_error() gives a thread-local message.
 *   - no CPU fallback and no multi-arch dispatch: everything requires an sm_100 device.
 *   - activations are channels-last: 1-D features [B, T, C], images [B, T, F, C].
 */
#ifndef ARTSPEECH_B200_H_
#define ARTSPEECH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum as_status {
  AS_OK = 0,
  AS_ERR_SHAPE = -1,
  AS_ERR_DTYPE = -2,
  AS_ERR_ALIGN = -3,
  AS_ERR_ARCH = -4,
  AS_ERR_CUDA = -5,
  AS_ERR_WORKSPACE = -6
} as_status;

typedef enum as_dtype { AS_F16 = 0, AS_BF16 = 1, AS_F32 = 2 } as_dtype;

typedef enum as_act {
  AS_ACT_NONE = 0,
  AS_ACT_LRELU = 1, /* x > 0 ? x : slope * x */
  AS_ACT_RELU = 2,
  AS_ACT_TANH = 3,
  AS_ACT_SWISH = 4, /* x * sigmoid(x) */
  AS_ACT_ABS = 5
} as_act;

int as_version(void);
const char* as_last_error(void);

/* ------------------------------------------------------------------------------------------
 * MAS — replaces S_monotonic_align.py:5-47 (maximum_path1), :50-95 (maximum_path2) and the
 * Triton kernel S_monotonic_align_Triton.py:7-71.
 *   value [B, Tx, Ty] fp32 row-major (not modified); x_len / y_len int32 [B];
 *   path  [B, Tx, Ty] fp32, fully overwritten with 0/1.
 *   tie_mode 0: stay on ties (maximum_path2 / Triton);  1: move to x-1 on ties (maximum_path1).
 *   workspace: as_mas_workspace_bytes() bytes (may be 0 → NULL allowed).
 * ------------------------------------------------------------------------------------------ */
size_t as_mas_workspace_bytes(int32_t B, int32_t Tx, int32_t Ty);
int as_mas_maximum_path(const float* value, const int32_t* x_len, const int32_t* y_len,
                        float* path, int32_t B, int32_t Tx, int32_t Ty, int32_t tie_mode,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces every dense torch Conv1d / ConvTranspose1d (polyphase-packed) / Conv2d / Linear on
 * the path: Vocoder/vocoder.py:35-42,100-116; models.py:189-202 (AdainResBlk1d convs),
 * :497-517 (Decoder), :596-621 (ArtsPredictor); Utils/RelTransformerEnc.py:127-136,261-269,
 * 316-325; Utils/JDC/model.py:102-137; Utils/EMA/conformer (Linear / pointwise convs).
 *
 *   acc[b,to,fo,co] = sum_{j<ntaps} sum_{ci<Cin} x[b, to+dt[j], fo+df[j], ci] * w[j][co][ci]
 *                     (out-of-range input coordinates read as zero: TMA OOB fill)
 *   v     = (acc + bias[co] + res1[row,co] + res2[row,co]) * out_scale,  row = (b*To+to)*Fo+fo
 *   v     = 0 where lens != NULL and to >= lens[b]
 *   y_raw[row*y_raw_ld + co] = v            (if y_raw)
 *   y_act[row*y_act_ld + co] = act(v)       (if y_act; act(0) = 0 for every supported act)
 *
 * x: 16-bit (AS_F16 or AS_BF16), rows of x_ld elements (x_ld % 8 == 0, base 16-byte aligned).
 * w: same dtype, packed [ntaps][CoutP][CinP], CinP % bk == 0 where bk = (Cin >= 64 ? 64 : 32),
 *    CoutP % tile_n == 0 (tile_n from as_conv_tile_n(Cout)); padding must be zero.
 * fp32 accumulation.  Output / residual dtypes: any of as_dtype.
 * ------------------------------------------------------------------------------------------ */
typedef struct as_conv_params {
  const void* x;
  int32_t x_dtype;
  int32_t B, T, F, Cin;
  int64_t x_ld;
  const void* w;
  int32_t ntaps, CinP, CoutP, Cout;
  const int32_t* tap_dt; /* host [ntaps] */
  const int32_t* tap_df; /* host [ntaps] */
  int32_t To, Fo;        /* output spatial size */
  const float* bias;     /* [Cout] fp32 or NULL */
  const void* res1;
  int32_t res1_dtype;
  int64_t res1_ld;
  const void* res2;
  int32_t res2_dtype;
  int64_t res2_ld;
  float out_scale;
  void* y_raw;
  int32_t y_raw_dtype;
  int64_t y_raw_ld;
  void* y_act;
  int32_t y_act_dtype;
  int64_t y_act_ld;
  int32_t act;
  float slope;
  const int32_t* lens; /* device [B] or NULL */
  /* optional fused instance-norm statistics of v (models.py:230-240): stats[b][co][0] += sum,
   * stats[b][co][1] += sum of squares over valid rows (fp32 atomics; caller zeroes it). */
  float* stats;
} as_conv_params;

int32_t as_conv_tile_n(int32_t Cout);
int as_conv_igemm(const as_conv_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARTSPEECH_B200_H_ */
