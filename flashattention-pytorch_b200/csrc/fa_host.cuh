// Host-side helpers shared by the C-ABI entry points: TMA tensor-map encoding through the driver entry point
// (no -lcuda link: the library must build and load on a GPU-less box), argument validation, error codes.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "../../include/fa_sm100.h"

namespace fa {

enum : int { kElemF16 = 0, kElemBF16 = 1, kElemF32 = 2, kElemU8 = 3 };

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

// 3-D tensor (cols, rows, slices) over a dense-row tensor; box = (box_cols, box_rows, 1).
// 128-byte swizzle when box_cols * elem_size == 128, which is what every UMMA operand tile here uses.
inline int make_tmap_3d(CUtensorMap* out, const void* ptr, int elem, uint64_t cols, uint64_t rows, uint64_t slices,
                        uint64_t slice_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  auto enc = tensor_map_encoder();
  if (!enc) return FA_SM100_EDRIVER;
  const uint64_t esz = (elem == kElemF32) ? 4 : (elem == kElemU8) ? 1 : 2;
  CUtensorMapDataType dt = elem == kElemF32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem == kElemBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : elem == kElemU8   ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t gdim[3] = {cols, rows, slices};
  cuuint64_t gstride[2] = {cols * esz, slice_stride_elems * esz};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = (box_cols * esz == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dt, 3, const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FA_SM100_OK : FA_SM100_EDRIVER;
}

inline bool aligned16(const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Geometry {
  long long bh, n_q, n_kv, q_bh_stride, kv_bh_stride, lse_bh_stride;
  int d, dp, dtype, causal, diag;  // diag = q_row0 - kv_col0: key c visible to query r iff c <= r + diag
                                   // dp = d rounded up to the kernel variant's head dim (64 / 128)
  float scale;
};

inline int check_shape(const fa_sm100_shape* s, Geometry* g, int max_d = 128) {
  if (!s) return FA_SM100_EINVAL_PTR;
  if (s->dtype != FA_SM100_DTYPE_F16 && s->dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  // any multiple of 8 up to 128 (256 for the plain forward): rows stay 16-byte aligned for TMA, and the tensor maps carry the true d, so the
  // columns between d and the kernel variant's 64 / 128 are zero-filled on load and clipped on store (no pad copies)
  if (s->d < 8 || s->d > max_d || (s->d % 8)) return FA_SM100_EINVAL_HEADDIM;
  if (s->bh <= 0 || s->n_q <= 0 || s->n_kv <= 0) return FA_SM100_EINVAL_SHAPE;
  if (s->n_q > (1ll << 30) || s->n_kv > (1ll << 30) || s->bh > (1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  if (!(s->softmax_scale > 0.f) || !std::isfinite(s->softmax_scale)) return FA_SM100_EINVAL_SCALE;
  const long long diag = s->q_row0 - s->kv_col0;
  if (diag > (1ll << 30) || diag < -(1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  g->bh = s->bh;
  g->n_q = s->n_q;
  g->n_kv = s->n_kv;
  g->d = s->d;
  g->dp = s->d <= 64 ? 64 : s->d <= 128 ? 128 : 256;
  g->dtype = s->dtype;
  g->causal = s->causal ? 1 : 0;
  g->diag = static_cast<int>(diag);
  g->scale = s->softmax_scale;
  g->q_bh_stride = s->q_bh_stride ? s->q_bh_stride : s->n_q * s->d;
  g->kv_bh_stride = s->kv_bh_stride ? s->kv_bh_stride : s->n_kv * s->d;
  g->lse_bh_stride = s->lse_bh_stride ? s->lse_bh_stride : s->n_q;
  if (g->q_bh_stride < s->n_q * s->d || g->kv_bh_stride < s->n_kv * s->d || g->lse_bh_stride < s->n_q)
    return FA_SM100_EINVAL_SHAPE;
  if ((g->q_bh_stride % 8) || (g->kv_bh_stride % 8)) return FA_SM100_EINVAL_SHAPE;  // TMA: 16-byte strides
  return FA_SM100_OK;
}

inline int check_device() {
  static int cached[64];  // 0 = unknown, 1 = sm_100, 2 = something else (benign race: idempotent writes)
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return FA_SM100_EDEVICE;
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev] == 1 ? FA_SM100_OK : FA_SM100_EDEVICE;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return FA_SM100_EDEVICE;
  if (dev >= 0 && dev < 64) cached[dev] = (major == 10) ? 1 : 2;
  return major == 10 ? FA_SM100_OK : FA_SM100_EDEVICE;
}

// log2 of the slices per scheduling group for cost-skewed (causal) grids: about two waves of CTAs per group (see the
// note on work-item order in ptx.cuh).  FA_SM100_SCHED_GROUP overrides the group size (tuning knob: 1 = slice-major order).
inline int sched_group_log2(bool skewed, long long n_ranks, long long n_slices) {
  static int forced = [] {
    const char* e = std::getenv("FA_SM100_SCHED_GROUP");
    return e ? std::atoi(e) : 0;
  }();
  long long g = 1;
  if (forced > 0) g = forced;
  else if (skewed) g = (2 * 148) / (n_ranks > 0 ? n_ranks : 1);
  int lg = 0;
  while ((2ll << lg) <= g && (2ll << lg) <= n_slices) ++lg;  // largest power of two <= min(g, n_slices)
  return lg;
}
// Validated form of fa_sm100_ext.
struct ExtArgs {
  const uint8_t* block_mask = nullptr;
  long long mask_bh_stride = 0;
  int mask_cols = 0;
  uint32_t seed_lo = 0, seed_hi = 0, rng_offset = 0, drop_threshold = 0;
  float drop_scale = 1.f;
  long long q_row0 = 0, kv_col0 = 0;
};
inline int check_ext(const fa_sm100_ext* e, const fa_sm100_shape* s, int max_tiles, ExtArgs* out) {
  *out = ExtArgs();
  out->q_row0 = s->q_row0;
  out->kv_col0 = s->kv_col0;
  out->mask_cols = static_cast<int>((s->n_kv + 127) / 128);
  if (e == nullptr) return FA_SM100_OK;
  if (!(e->dropout_p >= 0.f) || !(e->dropout_p < 1.f)) return FA_SM100_EINVAL_EXT;
  const uint32_t thr = static_cast<uint32_t>(e->dropout_p * 256.f);  // floor: p quantised to 1/256
  if (thr > 0) {
    if ((s->q_row0 % 4) || (s->kv_col0 % 4) || s->q_row0 < 0 || s->kv_col0 < 0) return FA_SM100_EINVAL_EXT;
    out->drop_threshold = thr;
    out->drop_scale = 256.f / static_cast<float>(256u - thr);
    out->seed_lo = static_cast<uint32_t>(e->seed);
    out->seed_hi = static_cast<uint32_t>(e->seed >> 32);
    out->rng_offset = static_cast<uint32_t>(e->offset);
  }
  if (e->block_mask != nullptr) {
    const long long rows = (s->n_q + 127) / 128, cols = (s->n_kv + 127) / 128;
    if (rows > max_tiles || cols > max_tiles) return FA_SM100_EINVAL_EXT;
    if (e->mask_bh_stride != 0 && e->mask_bh_stride < rows * cols) return FA_SM100_EINVAL_EXT;
    out->block_mask = e->block_mask;
    out->mask_bh_stride = e->mask_bh_stride;
  }
  return FA_SM100_OK;
}
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? FA_SM100_OK : FA_SM100_ELAUNCH; }

}  // namespace fa
