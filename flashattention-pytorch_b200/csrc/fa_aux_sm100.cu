// HBM-bound helper kernels of the backward (dQ finish, fp32 -> 16-bit cast) and library metadata.
#include "ptx.cuh"
#include "fa_host.cuh"

namespace fa {

// out = cast(acc * alpha), 8 elements per thread (two 16-byte loads, one 16-byte store)
template <bool kBF16>
__global__ void __launch_bounds__(256) fa_cast_scaled_kernel(const float* __restrict__ acc, uint16_t* __restrict__ out,
                                                             long long n, float alpha) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(acc + i);
      const float4 b = *reinterpret_cast<const float4*>(acc + i + 4);
      uint4 w;
      w.x = pack2<kBF16>(a.x * alpha, a.y * alpha);
      w.y = pack2<kBF16>(a.z * alpha, a.w * alpha);
      w.z = pack2<kBF16>(b.x * alpha, b.y * alpha);
      w.w = pack2<kBF16>(b.z * alpha, b.w * alpha);
      *reinterpret_cast<uint4*>(out + i) = w;
    } else {
      for (long long j = i; j < n; ++j) {
        const uint32_t w = pack2<kBF16>(acc[j] * alpha, 0.f);
        out[j] = static_cast<uint16_t>(w & 0xFFFFu);
      }
    }
  }
}

// strided variant for dQ: rows dense, slices strided
template <bool kBF16>
__global__ void __launch_bounds__(256) fa_dq_finish_kernel(const float* __restrict__ acc, uint16_t* __restrict__ out,
                                                           long long bh, long long slice_elems, long long q_bh_stride,
                                                           float alpha) {
  const long long per_slice_vec = slice_elems / 8;
  const long long total = bh * per_slice_vec;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const long long b = v / per_slice_vec, e = (v % per_slice_vec) * 8;
    const float* src = acc + b * q_bh_stride + e;  // the accumulator shares q's slice stride
    const float4 x = *reinterpret_cast<const float4*>(src);
    const float4 y = *reinterpret_cast<const float4*>(src + 4);
    uint4 w;
    w.x = pack2<kBF16>(x.x * alpha, x.y * alpha);
    w.y = pack2<kBF16>(x.z * alpha, x.w * alpha);
    w.z = pack2<kBF16>(y.x * alpha, y.y * alpha);
    w.w = pack2<kBF16>(y.z * alpha, y.w * alpha);
    *reinterpret_cast<uint4*>(out + b * q_bh_stride + e) = w;
  }
}

inline int grid_for(long long work_items, int per_block, int cap = 148 * 16) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}

}  // namespace fa

extern "C" int fa_sm100_version(void) { return FA_SM100_VERSION; }

extern "C" const char* fa_sm100_strerror(int code) {
  switch (code) {
    case FA_SM100_OK: return "ok";
    case FA_SM100_EINVAL_DTYPE: return "unsupported dtype (need fp16 or bf16)";
    case FA_SM100_EINVAL_HEADDIM: return "unsupported head dim (need a multiple of 8 in [8, 128])";
    case FA_SM100_EINVAL_SHAPE: return "invalid shape or stride";
    case FA_SM100_EINVAL_PTR: return "null or misaligned pointer";
    case FA_SM100_EINVAL_SCALE: return "softmax_scale must be finite and > 0";
    case FA_SM100_EDRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case FA_SM100_ELAUNCH: return "CUDA kernel launch failed";
    case FA_SM100_EDEVICE: return "current CUDA device is not sm_100";
    case FA_SM100_EINVAL_EXT: return "bad extras (dropout_p outside [0,1), offsets not multiples of 4, or too many tiles for the mask)";
    default: return "unknown fa_sm100 error";
  }
}

extern "C" size_t fa_sm100_dq_accum_bytes(const fa_sm100_shape* s) {
  if (!s || s->bh <= 0 || s->n_q <= 0 || s->d <= 0) return 0;
  return static_cast<size_t>(s->bh) * static_cast<size_t>(s->n_q) * static_cast<size_t>(s->d) * sizeof(float);
}

extern "C" int fa_sm100_dq_finish(const fa_sm100_shape* s, const float* dq_accum, void* dq, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g, /*max_d=*/256);
  if (rc) return rc;
  if (!fa::aligned16(dq_accum) || !fa::aligned16(dq)) return FA_SM100_EINVAL_PTR;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long slice = g.n_q * g.d;
  const int grid = fa::grid_for(g.bh * slice / 8, 256);
  if (g.dtype == FA_SM100_DTYPE_BF16)
    fa::fa_dq_finish_kernel<true><<<grid, 256, 0, st>>>(dq_accum, static_cast<uint16_t*>(dq), g.bh, slice, g.q_bh_stride, g.scale);
  else
    fa::fa_dq_finish_kernel<false><<<grid, 256, 0, st>>>(dq_accum, static_cast<uint16_t*>(dq), g.bh, slice, g.q_bh_stride, g.scale);
  return fa::launch_status();
}

extern "C" int fa_sm100_cast_scaled(const float* acc, void* out, int64_t n, float alpha, int32_t dtype,
                                    void* stream) {
  if (dtype != FA_SM100_DTYPE_F16 && dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  if (n <= 0) return FA_SM100_EINVAL_SHAPE;
  if (!fa::aligned16(acc) || !fa::aligned16(out)) return FA_SM100_EINVAL_PTR;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = fa::grid_for((n + 7) / 8, 256);
  if (dtype == FA_SM100_DTYPE_BF16)
    fa::fa_cast_scaled_kernel<true><<<grid, 256, 0, st>>>(acc, static_cast<uint16_t*>(out), n, alpha);
  else
    fa::fa_cast_scaled_kernel<false><<<grid, 256, 0, st>>>(acc, static_cast<uint16_t*>(out), n, alpha);
  return fa::launch_status();
}
