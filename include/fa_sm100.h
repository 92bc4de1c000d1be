/*
 * fa_sm100.h — C ABI of libfa_sm100.so: fused tiled attention forward + backward for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's native extension.  The reference binds six functions on one
 * pybind11 module (reference csrc/common/torch.extension.cpp:73-83):
 *     fa1_forward / forward / fa3_forward      -> fa_sm100_fwd
 *     fa1_backward / backward / fa3_backward   -> fa_sm100_bwd_prepare + fa_sm100_bwd + fa_sm100_dq_finish
 * FA1/FA2/FA3 are the same mathematical operator in the reference (csrc/fa1/fa1_fwd.cu:30-107,
 * csrc/fa2/fa2_fwd.cu:30-106, csrc/fa3/fa3_fwd.cu:103-211), so one kernel family serves all three.
 *
 * Conventions
 *   - plain C: pointers, sizes, a POD shape struct; no torch types, no exceptions, no allocation inside the library.
 *   - all tensor pointers are DEVICE pointers, 16-byte aligned; rows are dense (row stride == d elements).
 *     q/o/do/dq: (bh, n_q, d)   k/v/dk/dv: (bh, n_kv, d)   lse/delta: (bh, n_q) fp32.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls only enqueue work.
 *   - every function returns 0 on success or a negative FA_SM100_E* code; fa_sm100_strerror() names it.
 *   - re-entrant and thread-safe: no global mutable state except a once-initialised driver entry point.
 */
#ifndef FA_SM100_H_
#define FA_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_SM100_VERSION 100

#define FA_SM100_DTYPE_F16 0
#define FA_SM100_DTYPE_BF16 1
#define FA_SM100_DTYPE_F32 2 /* only for the *_f32 entry points */

#define FA_SM100_OK 0
#define FA_SM100_EINVAL_DTYPE (-1)   /* dtype is not fp16 / bf16 */
#define FA_SM100_EINVAL_HEADDIM (-2) /* d is not a multiple of 8 in [8, 256], or beyond what this form takes  */
#define FA_SM100_EINVAL_SHAPE (-3)   /* non-positive sizes, sizes beyond int32 tile indexing, bad strides */
#define FA_SM100_EINVAL_PTR (-4)     /* NULL or mis-aligned tensor pointer */
#define FA_SM100_EINVAL_SCALE (-5)   /* softmax_scale must be finite and > 0 */
#define FA_SM100_EDRIVER (-6)        /* could not resolve / call cuTensorMapEncodeTiled */
#define FA_SM100_ELAUNCH (-7)        /* kernel launch failed (cudaGetLastError) */
#define FA_SM100_EDEVICE (-8)        /* current device is not compute capability 10.x */
#define FA_SM100_EINVAL_EXT (-9)     /* bad fa_sm100_ext: dropout_p outside [0,1), offsets not multiples of 4, too many tiles */

/* Problem geometry shared by all entry points. */
typedef struct fa_sm100_shape {
  int64_t bh;            /* number of independent (batch*head) slices                                  */
  int64_t n_q;           /* query rows per slice                                                       */
  int64_t n_kv;          /* key/value rows per slice                                                   */
  int32_t d;             /* head dim: any multiple of 8 up to 256 (the *_ex, *_accum, *_f32 and fp8    */
                         /* forms: up to 128 / exactly 128, see each entry)                            */
  int32_t dtype;         /* FA_SM100_DTYPE_*                                                           */
  int32_t causal;        /* 0/1; key c is visible to query r iff kv_col0 + c <= q_row0 + r             */
  float softmax_scale;   /* S = Q K^T * softmax_scale                                                  */
  int64_t q_row0;        /* global sequence index of query row 0 (ring attention; 0 otherwise)         */
  int64_t kv_col0;       /* global sequence index of key row 0                                         */
  int64_t q_bh_stride;   /* elements between slices of q / o / do / dq / dq_accum (0 => n_q * d)       */
  int64_t kv_bh_stride;  /* elements between slices of k / v / dk / dv                (0 => n_kv * d)  */
  int64_t lse_bh_stride; /* elements between slices of lse / delta                    (0 => n_q)       */
} fa_sm100_shape;

/*
 * Optional extras of the *_ex entry points (the block-sparse mask and dropout of the
 * reference's stand-alone module, src/fa3/torch/flashattention_pytorch.py:80-87,94-174):
 *   block_mask : uint8 DEVICE array, (ceil(n_q/128) x ceil(n_kv/128)) row-major per slice, or one shared by all slices
 *                (mask_bh_stride = 0).  Tile (i, j) is computed iff block_mask[i][j] != 0; a skipped tile contributes
 *                nothing (as if its scores were -inf).  NULL = dense.  Block size is fixed at 128 (the reference's
 *                default, :19); the causal mask, if any, applies on top.
 *   dropout    : probabilities are dropped after the softmax: O = (P o keep / (1 - p)) V; lse is unaffected.
 *                keep bits come from Philox4x32-7 keyed by (seed, offset) and the element's GLOBAL (slice, query,
 *                key) coordinates, so forward and backward regenerate the same bits and a sharded run equals the
 *                unsharded one.  p is quantised to multiples of 1/256: an element is dropped iff its random byte <
 *                floor(p * 256).  q_row0 and kv_col0 must be multiples of 4 when dropout is on.
 */
typedef struct fa_sm100_ext {
  const uint8_t* block_mask;
  int64_t mask_bh_stride; /* elements between the masks of consecutive slices; 0 = one mask shared by all slices */
  float dropout_p;        /* [0, 1); 0 = no dropout */
  uint64_t seed;
  uint64_t offset;
} fa_sm100_ext;

int fa_sm100_version(void);
const char* fa_sm100_strerror(int code);

/* Scratch the caller must provide to the backward:
 *   dq_accum : fp32 dQ accumulator with q's geometry (bh slices of n_q * d, slice stride q_bh_stride), ZEROED before
 *              the first fa_sm100_bwd that adds into it (by the caller, or by fa_sm100_bwd_prepare on request);
 *   rowstats : per-query-row statistics packed per 128-row tile, bh * ceil(n_q/128) * 256 floats. */
size_t fa_sm100_dq_accum_bytes(const fa_sm100_shape* s);
size_t fa_sm100_rowstats_bytes(const fa_sm100_shape* s);

/*
 * Forward.  Replaces fa{1,2,3}_forward (reference csrc/fa1/fa1_fwd.cu:30-107):
 *   O = softmax(Q K^T * scale [+ causal mask]) V      -> o   (q's dtype)
 *   lse = rowmax + log(rowsum), natural log units     -> lse (fp32)
 * Rows with no visible key produce O = 0, lse = -inf.
 * Ring attention: if o_prev / lse_prev are non-NULL the new partial is merged with them by log-sum-exp
 *   lse' = logaddexp(lse_prev, lse);  O' = e^{lse_prev-lse'} O_prev + e^{lse-lse'} O
 * before write-out (o_prev may alias o, lse_prev may alias lse).
 */
int fa_sm100_fwd(const fa_sm100_shape* s, const void* q, const void* k, const void* v, void* o, float* lse,
                 const void* o_prev, const float* lse_prev, void* stream);

/* Forward with the extras above (ext == NULL behaves like fa_sm100_fwd without a previous partial). */
int fa_sm100_fwd_ex(const fa_sm100_shape* s, const fa_sm100_ext* ext, const void* q, const void* k, const void* v,
                    void* o, float* lse, void* stream);

/*
 * Backward pre-pass (reference csrc/fa1/fa1_bwd.cu:57): delta[r] = sum_c dO[r,c] * O[r,c], packed together with
 * lse[r] * log2(e) into `rowstats`, both NEGATED: for slice b and 128-row tile t, floats [(b*T + t)*256, +128) hold
 * -lse*log2e and the next 128 hold -delta (T = ceil(n_q/128)); rows past n_q get -inf / 0 so they contribute nothing.
 * `lse` is the (bh, n_q) tensor the forward returned (lse_bh_stride applies).
 * If `dq_accum_zero` is non-NULL the same launch also zero-fills that fp32 dQ accumulator (q's geometry), so the
 * caller needs no separate memset before fa_sm100_bwd.
 */
int fa_sm100_bwd_prepare(const fa_sm100_shape* s, const void* o, const void* d_o, const float* lse, float* rowstats,
                         float* dq_accum_zero, void* stream);

/*
 * Backward main pass, KV-outer (reference csrc/fa1/fa1_bwd.cu:70-110 with the Python skip rule of
 * src/fa1/torch/impl.py:89): recomputes P = exp(S - lse), then
 *   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  dK = scale * dS^T Q   -> dk, dv (input dtype)
 *   dQ partials (unscaled) are reduce-added in fp32 into dq_accum (see above).
 */
int fa_sm100_bwd(const fa_sm100_shape* s, const void* q, const void* k, const void* v, const void* d_o,
                 const float* rowstats, float* dq_accum, void* dk, void* dv, void* stream);

/* Backward main pass with the extras above: the same mask / (p, seed, offset) the forward ran with.  With dropout,
 * dV = (P o keep / (1-p))^T dO,  dP = (dO V^T) o keep / (1-p),  dS = P o (dP - delta)  (delta from the dropped O). */
int fa_sm100_bwd_ex(const fa_sm100_shape* s, const fa_sm100_ext* ext, const void* q, const void* k, const void* v,
                    const void* d_o, const float* rowstats, float* dq_accum, void* dk, void* dv, void* stream);

/*
 * Ring-attention form of the main pass: instead of writing 16-bit dk / dv, the fp32 partials (dK already scaled) go to
 * `dk_accum` / `dv_accum`, fp32 (bh, n_kv, d) with slice stride `acc_bh_stride` elements (0 => n_kv * d): reduce-added
 * (overwrite = 0) or stored (overwrite = 1: every element is written, zeros where a K/V row saw no query).  No 16-bit
 * rounding of partials; the ring schedule adds them to the fp32 accumulators that travel with the K/V block.
 */
int fa_sm100_bwd_accum(const fa_sm100_shape* s, const void* q, const void* k, const void* v, const void* d_o,
                       const float* rowstats, float* dq_accum, float* dk_accum, float* dv_accum,
                       int64_t acc_bh_stride, int32_t overwrite, void* stream);

/*
 * fp32 inputs (BASELINE config C1; the reference up-casts everything to fp32: csrc/fa1/fa1_fwd.cu:67,79-80, and its
 * tests hold fp32 to rtol = atol = 1e-4: tests/utils.py:31-36).  These two entry points keep that contract with plain
 * fp32 FMA arithmetic on the CUDA cores (no tensor-core rounding): same operator, same layouts with fp32 elements,
 * shape.dtype = FA_SM100_DTYPE_F32, d a multiple of 4 up to 128, bh <= 65535.  The backward is self-contained:
 * `delta_ws` is a scratch of bh * n_q floats, dq / dk / dv are written in fp32 (dq is zero-filled, then accumulated
 * with fp32 atomics, so its last bits vary from run to run).
 */
int fa_sm100_fwd_f32(const fa_sm100_shape* s, const float* q, const float* k, const float* v, float* o, float* lse,
                     void* stream);
int fa_sm100_bwd_f32(const fa_sm100_shape* s, const float* q, const float* k, const float* v, const float* o,
                     const float* d_o, const float* lse, float* delta_ws, float* dq, float* dk, float* dv, void* stream);

/*
 * FP8 forward (the path behind the reference's fp8=True flag, src/fa3/op.py:7,
 * src/fa3/cuda/impl.py:40-55, whose torch emulation is src/fa3/torch/impl.py:20-72,123-131).  Head dim 128 only.
 *   fa_sm100_fp8_quantize: x (bh, n, 128) fp16/bf16 -> out8 (bh, n, 128) e4m3 bytes (dense) + one fp32 scale per
 *     128-row block in scales[bh * ceil(n/128)]; with `hadamard` the rows are first sign-flipped (Philox bits of `seed`),
 *     Walsh-Hadamard transformed and divided by sqrt(128) -- apply it to Q and K with the SAME seed, not to V.
 *   fa_sm100_fwd_fp8: both products on `tcgen05.mma kind::f8f6f4`; o in shape.dtype, lse fp32.  sv_ref[bh] must be
 *     the largest V scale of the slice (V's per-block scale is folded into the e4m3 probabilities relative to it).
 */
int fa_sm100_fp8_quantize(const void* x, void* out8, float* scales, int64_t bh, int64_t n, int32_t d,
                          int64_t x_bh_stride, int32_t dtype, int32_t hadamard, uint64_t seed, void* stream);
int fa_sm100_fwd_fp8(const fa_sm100_shape* s, const void* q8, const void* k8, const void* v8, const float* sq,
                     const float* sk, const float* sv, const float* sv_ref, void* o, float* lse, void* stream);

/* dq[i] = cast(dq_accum[i] * softmax_scale).  (The reference scales per tile: csrc/fa1/fa1_bwd.cu:102-103.) */
int fa_sm100_dq_finish(const fa_sm100_shape* s, const float* dq_accum, void* dq, void* stream);

/* out[i] = cast(acc[i] * alpha) for n contiguous elements: converts fp32 ring accumulators to `dtype`. */
int fa_sm100_cast_scaled(const float* acc, void* out, int64_t n, float alpha, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FA_SM100_H_ */
