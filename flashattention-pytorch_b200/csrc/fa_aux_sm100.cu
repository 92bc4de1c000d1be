// HBM-bound helper kernels of the backward (dQ finish, fp32 -> 16-bit cast), library metadata, and
// the UMMA/TMA descriptor bring-up probe.
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

// ------------------------------------------------------------------------------------------------
// UMMA probe: one CTA, one 128x128x128 product through each operand path.
// ------------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(128, 1)
fa_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const uint16_t* __restrict__ a_gmem, float* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;           // 2 sub-tiles of [128 rows][128 B]
  uint8_t* b_smem = smem + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_d = tmem_base;        // D accumulator: columns [0,128)
  const uint32_t t_a = tmem_base + 128;  // A operand (mode 2): columns [128,192)

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 65536);
    for (int c = 0; c < 2; ++c) {
      tma_load_3d(a_smem + c * 16384, &tm_a, &bars[0], c * 64, 0, 0);
      tma_load_3d(b_smem + c * 16384, &tm_b, &bars[0], c * 64, 0, 0);
    }
  }
  if (mode == 2) {
    // thread r packs row r of A into TMEM (two 16-bit values per 32-bit column)
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a_gmem + threadIdx.x * 128);
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) w[x] = arow[q * 16 + x];
      tmem_st16(t_a + lane_sel + q * 16, w);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t koff_kmajor = (kk >> 2) * 16384 + (kk & 3) * 32;  // 16 K-elements inside a 128-B row
      const uint32_t koff_mnmajor = kk * 16 * 128;                     // 16 K-rows of 128 B
      const uint32_t acc = kk > 0 ? 1u : 0u;
      if (mode == 0) {
        umma_ss(t_d, umma_smem_desc(a0 + koff_kmajor, 16, 1024), umma_smem_desc(b0 + koff_kmajor, 16, 1024),
                umma_idesc(kBF16, 128, 128, false, false), acc);
      } else if (mode == 1) {
        umma_ss(t_d, umma_smem_desc(a0 + koff_kmajor, 16, 1024), umma_smem_desc(b0 + koff_mnmajor, 16384, 1024),
                umma_idesc(kBF16, 128, 128, false, true), acc);
      } else if (mode == 2) {
        umma_ts(t_d, t_a + kk * 8, umma_smem_desc(b0 + koff_mnmajor, 16384, 1024),
                umma_idesc(kBF16, 128, 128, false, true), acc);
      } else {
        umma_ss(t_d, umma_smem_desc(a0 + koff_mnmajor, 16384, 1024), umma_smem_desc(b0 + koff_mnmajor, 16384, 1024),
                umma_idesc(kBF16, 128, 128, true, true), acc);
      }
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float* orow = out + threadIdx.x * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[32];
      tmem_ld32(t_d + lane_sel + q * 32, reinterpret_cast<uint32_t*>(v));
      tc_wait_ld();
#pragma unroll
      for (int x = 0; x < 32; ++x) orow[q * 32 + x] = v[x];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
  (void)lane;
}

}  // namespace fa

extern "C" int fa_sm100_version(void) { return FA_SM100_VERSION; }

extern "C" const char* fa_sm100_strerror(int code) {
  switch (code) {
    case FA_SM100_OK: return "ok";
    case FA_SM100_EINVAL_DTYPE: return "unsupported dtype (need fp16 or bf16)";
    case FA_SM100_EINVAL_HEADDIM: return "unsupported head dim (need 64 or 128; pad other sizes)";
    case FA_SM100_EINVAL_SHAPE: return "invalid shape or stride";
    case FA_SM100_EINVAL_PTR: return "null or misaligned pointer";
    case FA_SM100_EINVAL_SCALE: return "softmax_scale must be finite and > 0";
    case FA_SM100_EDRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case FA_SM100_ELAUNCH: return "CUDA kernel launch failed";
    case FA_SM100_EDEVICE: return "current CUDA device is not sm_100";
    default: return "unknown fa_sm100 error";
  }
}

extern "C" size_t fa_sm100_dq_accum_bytes(const fa_sm100_shape* s) {
  if (!s || s->bh <= 0 || s->n_q <= 0 || s->d <= 0) return 0;
  return static_cast<size_t>(s->bh) * static_cast<size_t>(s->n_q) * static_cast<size_t>(s->d) * sizeof(float);
}

extern "C" int fa_sm100_dq_finish(const fa_sm100_shape* s, const float* dq_accum, void* dq, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g);
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

extern "C" int fa_sm100_probe_umma(int mode, int32_t dtype, const void* a, const void* b, float* out, void* stream) {
  if (dtype != FA_SM100_DTYPE_F16 && dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  if (mode < 0 || mode > 3) return FA_SM100_EINVAL_SHAPE;
  if (!fa::aligned16(a) || !fa::aligned16(b) || !fa::aligned16(out)) return FA_SM100_EINVAL_PTR;
  int rc = fa::check_device();
  if (rc) return rc;
  const int elem = dtype == FA_SM100_DTYPE_BF16 ? fa::kElemBF16 : fa::kElemF16;
  CUtensorMap tm_a, tm_b;
  if ((rc = fa::make_tmap_3d(&tm_a, a, elem, 128, 128, 1, 128 * 128, 64, 128))) return rc;
  if ((rc = fa::make_tmap_3d(&tm_b, b, elem, 128, 128, 1, 128 * 128, 64, 128))) return rc;
  const int smem = 65536 + 1024 + 64;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FA_SM100_DTYPE_BF16) {
    cudaFuncSetAttribute(fa::fa_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fa::fa_probe_kernel<true><<<1, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
  } else {
    cudaFuncSetAttribute(fa::fa_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fa::fa_probe_kernel<false><<<1, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
  }
  return fa::launch_status();
}
