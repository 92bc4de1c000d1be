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

enum : int { kElemF16 = 0, kElemBF16 = 1, kElemF32 = 2 };

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
  const uint64_t esz = (elem == kElemF32) ? 4 : 2;
  CUtensorMapDataType dt = elem == kElemF32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem == kElemBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
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

inline int check_shape(const fa_sm100_shape* s, Geometry* g) {
  if (!s) return FA_SM100_EINVAL_PTR;
  if (s->dtype != FA_SM100_DTYPE_F16 && s->dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  // any multiple of 8 up to 128: rows stay 16-byte aligned for TMA, and the tensor maps carry the true d, so the
  // columns between d and the kernel variant's 64 / 128 are zero-filled on load and clipped on store (no pad copies)
  if (s->d < 8 || s->d > 128 || (s->d % 8)) return FA_SM100_EINVAL_HEADDIM;
  if (s->bh <= 0 || s->n_q <= 0 || s->n_kv <= 0) return FA_SM100_EINVAL_SHAPE;
  if (s->n_q > (1ll << 30) || s->n_kv > (1ll << 30) || s->bh > (1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  if (!(s->softmax_scale > 0.f) || !std::isfinite(s->softmax_scale)) return FA_SM100_EINVAL_SCALE;
  const long long diag = s->q_row0 - s->kv_col0;
  if (diag > (1ll << 30) || diag < -(1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  g->bh = s->bh;
  g->n_q = s->n_q;
  g->n_kv = s->n_kv;
  g->d = s->d;
  g->dp = s->d <= 64 ? 64 : 128;
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
// Persistent grids: one CTA per SM, minus `sm_margin` SMs left free for concurrently running communication kernels
// (NCCL send/recv in the ring-attention schedule cannot make progress under 148 resident one-CTA-per-SM kernels).
// The margin is process-wide state set through fa_sm100_set_sm_margin() or the FA_SM100_SM_MARGIN environment variable.
inline int& sm_margin_ref() {
  static int margin = [] {
    const char* e = std::getenv("FA_SM100_SM_MARGIN");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 0;
  }();
  return margin;
}
inline long long persistent_ctas(long long n_items) {
  static const bool one_item_per_cta = [] {  // FA_SM100_PERSISTENT=0: measurement knob, one CTA per work item
    const char* e = std::getenv("FA_SM100_PERSISTENT");
    return e != nullptr && e[0] == '0';
  }();
  if (one_item_per_cta) return n_items;
  static int sms[64];  // per device; benign race: idempotent writes
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  int n = (dev >= 0 && dev < 64) ? sms[dev] : 0;
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return -1;
    if (dev >= 0 && dev < 64) sms[dev] = n;
  }
  long long ctas = n - sm_margin_ref();
  if (ctas < 1) ctas = 1;
  return n_items < ctas ? n_items : ctas;
}
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? FA_SM100_OK : FA_SM100_ELAUNCH; }

}  // namespace fa
