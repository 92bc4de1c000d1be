// Measurement aids, built into a SEPARATE debug library (libfa_sm100_probes.so, C ABI in include/fa_sm100_probes.h):
// the UMMA/TMA descriptor bring-up probe (single CTA and CTA pair), the tensor-core issue-rate probe and the L2
// reduce-add rate probe.  Nothing here is part of the product ABI or of the reference's surface.
#include "ptx.cuh"
#include "fa_host.cuh"
#include "../../include/fa_sm100_probes.h"

namespace fa {

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

// e4m3 probe (modes 6, 7): one 128x128x128 kind::f8f6f4 product.  mode 6: D = A B^T (A, B K-major in smem: rows of 128
// bytes).  mode 7: D = A B with A read from TMEM (four e4m3 per 32-bit column) and B MN-major in smem (row = k index).
__global__ void __launch_bounds__(128, 1)
fa_probe_fp8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const uint8_t* __restrict__ a_gmem, float* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;           // [128 rows][128 B]
  uint8_t* b_smem = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
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
  const uint32_t t_d = tmem_base, t_a = tmem_base + 128;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 32768);
    tma_load_3d(a_smem, &tm_a, &bars[0], 0, 0, 0);
    tma_load_3d(b_smem, &tm_b, &bars[0], 0, 0, 0);
  }
  if (mode == 7) {  // thread r packs row r of A into TMEM: 128 bytes = 32 columns
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a_gmem + threadIdx.x * 128);
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t w[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) w[x] = arow[x];
    tmem_st32(t_a + lane_sel, w);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
    for (int kk = 0; kk < 4; ++kk) {  // 32 K-elements per instruction
      const uint32_t acc = kk > 0 ? 1u : 0u;
      if (mode == 6) {
        umma_ss_f8(t_d, umma_smem_desc(a0 + kk * 32, 16, 1024), umma_smem_desc(b0 + kk * 32, 16, 1024),
                   umma_idesc(false, 128, 128, false, false), acc);
      } else {
        umma_ts_f8(t_d, t_a + kk * 8, umma_smem_desc(b0 + kk * 32 * 128, 16384, 1024),
                   umma_idesc(false, 128, 128, false, true), acc);
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
}

// CTA-pair probe (modes 4, 5): D[256x128] through one cta_group::2 product, the two operand paths a paired forward
// needs.  mode 4: A, B K-major from smem (S = Q K^T: each CTA stages 128 rows of A and 64 rows of B).
// mode 5: A from TMEM, B MN-major (O = P V: each CTA stages all 128 k-rows of its 64 output columns of B).
template <bool kBF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
fa_probe_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const uint16_t* __restrict__ a_gmem, float* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;           // this CTA's 128 rows of A: 2 sub-tiles of [128 rows][128 B]
  uint8_t* b_smem = smem + 32768;   // this CTA's half of B (16 KiB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);  // operands landed (used in the leader only)
    mbar_init(&bars[1], 1);  // product complete (one per CTA, multicast commit)
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_2sm(tmem_slot, 256);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_d = tmem_base;
  const uint32_t t_a = tmem_base + 128;

  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(&bars[0], mode == 4 ? 2u * (32768u + 16384u) : 2u * 16384u);
    const uint32_t lead_bar = mapa_shared(smem_u32(&bars[0]), 0);
    if (mode == 4) {
      for (int c = 0; c < 2; ++c) {
        tma_load_3d_2sm(a_smem + c * 16384, &tm_a, lead_bar, c * 64, static_cast<int>(rank) * 128, 0);
        tma_load_3d_2sm(b_smem + c * 8192, &tm_b, lead_bar, c * 64, static_cast<int>(rank) * 64, 0);
      }
    } else {
      tma_load_3d_2sm(b_smem, &tm_b, lead_bar, static_cast<int>(rank) * 64, 0, 0);
    }
  }
  if (mode == 5) {
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a_gmem + (rank * 128 + threadIdx.x) * 128);
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
  cluster_sync();
  tc_fence_after();

  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t acc = kk > 0 ? 1u : 0u;
      if (mode == 4) {
        umma_ss_2sm(t_d, umma_smem_desc(a0 + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                    umma_smem_desc(b0 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024),
                    umma_idesc(kBF16, 256, 128, false, false), acc);
      } else {
        umma_ts_2sm(t_d, t_a + kk * 8, umma_smem_desc(b0 + kk * 2048, 16384, 1024),
                    umma_idesc(kBF16, 256, 128, false, true), acc);
      }
    }
    tc_commit_2sm(&bars[1], 3);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float* orow = out + (rank * 128 + threadIdx.x) * 128;
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
  cluster_sync();  // the peer's smem and TMEM stay alive until both CTAs are done with the product
  if (warp == 0) tmem_dealloc_2sm(tmem_base, 256);
}

// MMA issue-rate probe: every CTA (or CTA pair) streams `groups` full K=128 products (8 UMMAs each) from fixed smem /
// TMEM operands into two alternating TMEM accumulators.  Nothing is read back; the host times the launch.  It answers
// "what does one S or PV product cost per SM in each operand configuration" without the rest of the attention loop.
template <bool kPair, bool kTS>
__global__ void __launch_bounds__(128, 1)
fa_mma_rate_kernel(int n, int groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;          // 128 rows x 128 K (2 sub-tiles of 16 KiB)
  uint8_t* b_smem = smem + 32768;  // up to 256 rows x 128 K
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;

  // bounded pseudo-random bf16/fp16 bit patterns (exponent field kept small: no inf/nan, realistic toggling)
  for (uint32_t i = threadIdx.x; i < 98304 / 4; i += 128) {
    uint32_t h = (i + 1u) * 2654435761u;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807F807Fu) | 0x3C003C00u;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (kPair) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (kTS) {
    uint32_t w[16];
#pragma unroll
    for (int x = 0; x < 16; ++x) w[x] = 0x3C003C00u + threadIdx.x + x;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_st16(tmem_base + 448 + lane_sel + q * 16, w);
    tc_wait_st();
    tc_fence_before();
    if (kPair) cluster_sync(); else __syncthreads();
    tc_fence_after();
  }

  const int n_local = kPair ? n / 2 : n;  // rows (K-major) or columns (MN-major) of B staged per CTA
  const uint32_t idesc = umma_idesc(true, kPair ? 256 : 128, n, false, kTS);
  const uint32_t d_stride = n > 128 ? 256u : 128u;
  if (warp == 0 && rank == 0) {
    const uint32_t a_lo = umma_desc_lo(smem_u32(a_smem), 16);
    const uint32_t b_lo = umma_desc_lo(smem_u32(b_smem), kTS ? 16384 : 16);
    const uint32_t b_sub = static_cast<uint32_t>(n_local) * 128u;  // K-major: next 64 K-columns of B
    for (int g = 0; g < groups; ++g) {
      const uint32_t t_d = tmem_base + (g & 1) * d_stride;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t acc = kk > 0 ? 1u : 0u;
          const uint32_t a_off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if (kTS) {
            const uint64_t bd = umma_desc(b_lo + ((kk * 2048) >> 4));
            if (kPair) umma_ts_2sm(t_d, tmem_base + 448 + kk * 8, bd, idesc, acc);
            else umma_ts(t_d, tmem_base + 448 + kk * 8, bd, idesc, acc);
          } else {
            const uint64_t bd = umma_desc(b_lo + (((kk >> 2) * b_sub + (kk & 3) * 32) >> 4));
            if (kPair) umma_ss_2sm(t_d, umma_desc(a_lo + a_off), bd, idesc, acc);
            else umma_ss(t_d, umma_desc(a_lo + a_off), bd, idesc, acc);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (kPair) tc_commit_2sm(&bars[0], 3); else tc_commit(&bars[0]);
    }
    __syncwarp();
  }
  if (warp == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
  }
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  if (warp == 0) {
    if (kPair) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// L2 reduce-add rate probe: the backward's dQ traffic pattern with nothing else going on.  CTA (slice, j) walks the
// query tiles of its slice (optionally starting at a rotated position) and TMA-reduce-adds a 128 x 128 fp32 tile
// (four 128 x 32 boxes from two alternating pairs of staging buffers) into `acc` for each of them.
__global__ void __launch_bounds__(128, 1)
fa_reduce_rate_kernel(const __grid_constant__ CUtensorMap tm_acc, float* __restrict__ acc, int nqt, int nkt,
                      int flags) {
  extern __shared__ __align__(1024) uint8_t smem_red[];
  const int slice = blockIdx.x / nkt, j = blockIdx.x % nkt;
  const bool rotate = flags & 1, from_regs = flags & 2;
  if (from_regs) {
    // register path: thread = query row (as after a TMEM load).  Lane pairs split each 32-byte sector of a row between
    // them, so one warp-wide red.v4 covers 16 rows x 32 contiguous bytes: full sectors, no shared-memory staging.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_a = warp * 32 + (lane & ~1), row_b = row_a + 1, sub = (lane & 1) * 4;
    for (int it = 0; it < nqt; ++it) {
      int i = it + (rotate ? j : 0);
      if (i >= nqt) i -= nqt;
      float* base = acc + (static_cast<size_t>(slice) * nqt + i) * 128 * 128;
#pragma unroll 4
      for (int m = 0; m < 16; ++m) {
        red_add_v4(base + row_a * 128 + m * 8 + sub, 1.f, 1.f, 1.f, 1.f);
        red_add_v4(base + row_b * 128 + m * 8 + sub, 1.f, 1.f, 1.f, 1.f);
      }
    }
    return;
  }
  for (uint32_t i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<float*>(smem_red)[i] = 1.0f;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = 0; it < nqt; ++it) {
      int i = it + (rotate ? j : 0);
      if (i >= nqt) i -= nqt;
      uint8_t* stage = smem_red + (it & 1) * 32768;
      tma_store_wait_read<1>();  // the pair of buffers used two tiles ago has been read
      tma_reduce_add_3d(&tm_acc, stage, 0, i * 128, slice);
      tma_reduce_add_3d(&tm_acc, stage + 16384, 32, i * 128, slice);
      tma_reduce_add_3d(&tm_acc, stage, 64, i * 128, slice);
      tma_reduce_add_3d(&tm_acc, stage + 16384, 96, i * 128, slice);
      tma_store_commit();
    }
    tma_store_wait_all<0>();
  }
}

// MUFU rate probe: every thread runs `iters` rounds of 8 independent exp2 chains in one of three forms --
// mode 0: ex2.approx.ftz.f32 (one result per MUFU op), 1: ex2.approx.ftz.f16x2, 2: ex2.approx.ftz.bf16x2 (two results per
// op).  The host times the launch: results per second = threads * iters * 8 * (1 or 2) / time.
__global__ void __launch_bounds__(256) fa_ex2_rate_kernel(int mode, int iters, float* __restrict__ sink) {
  float acc = 0.f;
  if (mode == 0) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * static_cast<float>(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] - 1.0f;  // keep the argument in range, on the FMA pipe
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i];
  } else {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = mode == 1 ? 0xB800B800u + threadIdx.x + i : 0xBF00BF00u + threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
      if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] ^= 0x80008000u;  // flip the signs back to negative arguments (ALU pipe)
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __uint_as_float(x[i] & 0x3FFFFFFFu);
  }
  if (acc == 123.456f) sink[0] = acc;  // never true: keeps the chains alive
}

}  // namespace fa

extern "C" int fa_sm100_probe_ex2_rate(int mode, int iters, int ctas, float* sink, void* stream) {
  if (mode < 0 || mode > 2 || iters <= 0 || ctas <= 0 || sink == nullptr) return FA_SM100_EINVAL_SHAPE;
  int rc = fa::check_device();
  if (rc) return rc;
  fa::fa_ex2_rate_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(mode, iters, sink);
  return fa::launch_status();
}

extern "C" int fa_sm100_probe_umma(int mode, int32_t dtype, const void* a, const void* b, float* out, void* stream) {
  if (mode == 6 || mode == 7) {  // e4m3 operands: a, b are 128 x 128 bytes
    if (!fa::aligned16(a) || !fa::aligned16(b) || !fa::aligned16(out)) return FA_SM100_EINVAL_PTR;
    int rc8 = fa::check_device();
    if (rc8) return rc8;
    CUtensorMap ta, tb;
    if ((rc8 = fa::make_tmap_3d(&ta, a, fa::kElemU8, 128, 128, 1, 128 * 128, 128, 128))) return rc8;
    if ((rc8 = fa::make_tmap_3d(&tb, b, fa::kElemU8, 128, 128, 1, 128 * 128, 128, 128))) return rc8;
    const int smem8 = 32768 + 1024 + 64;
    cudaFuncSetAttribute(fa::fa_probe_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8);
    fa::fa_probe_fp8_kernel<<<1, 128, smem8, static_cast<cudaStream_t>(stream)>>>(ta, tb, static_cast<const uint8_t*>(a),
                                                                                 out, mode);
    return fa::launch_status();
  }
  if (dtype != FA_SM100_DTYPE_F16 && dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  if (mode < 0 || mode > 5) return FA_SM100_EINVAL_SHAPE;
  if (!fa::aligned16(a) || !fa::aligned16(b) || !fa::aligned16(out)) return FA_SM100_EINVAL_PTR;
  int rc = fa::check_device();
  if (rc) return rc;
  const int elem = dtype == FA_SM100_DTYPE_BF16 ? fa::kElemBF16 : fa::kElemF16;
  const bool pair = mode >= 4;  // CTA-pair modes: A and out have 256 rows
  CUtensorMap tm_a, tm_b;
  if ((rc = fa::make_tmap_3d(&tm_a, a, elem, 128, pair ? 256 : 128, 1, (pair ? 256 : 128) * 128, 64, 128))) return rc;
  if ((rc = fa::make_tmap_3d(&tm_b, b, elem, 128, 128, 1, 128 * 128, 64, mode == 4 ? 64 : 128))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pair) {
    const int smem = 49152 + 1024 + 64;
    if (dtype == FA_SM100_DTYPE_BF16) {
      cudaFuncSetAttribute(fa::fa_probe_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      fa::fa_probe_pair_kernel<true><<<2, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
    } else {
      cudaFuncSetAttribute(fa::fa_probe_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      fa::fa_probe_pair_kernel<false><<<2, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
    }
    return fa::launch_status();
  }
  const int smem = 65536 + 1024 + 64;
  if (dtype == FA_SM100_DTYPE_BF16) {
    cudaFuncSetAttribute(fa::fa_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fa::fa_probe_kernel<true><<<1, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
  } else {
    cudaFuncSetAttribute(fa::fa_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fa::fa_probe_kernel<false><<<1, 128, smem, st>>>(tm_a, tm_b, static_cast<const uint16_t*>(a), out, mode);
  }
  return fa::launch_status();
}

template <bool kPair, bool kTS>
static int launch_mma_rate(int n, int groups, int ctas, cudaStream_t st) {
  const int smem = 98304 + 1024 + 64;
  cudaFuncSetAttribute(fa::fa_mma_rate_kernel<kPair, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, fa::fa_mma_rate_kernel<kPair, kTS>, n, groups) != cudaSuccess) {
    cudaGetLastError();
    return FA_SM100_ELAUNCH;
  }
  return fa::launch_status();
}

extern "C" int fa_sm100_probe_mma_rate(int pair, int a_from_tmem, int n, int groups, int ctas, void* stream) {
  if (n < 32 || n > 256 || (n % 32) || groups <= 0 || ctas <= 0 || (pair && (ctas & 1))) return FA_SM100_EINVAL_SHAPE;
  if (a_from_tmem && n > 128) return FA_SM100_EINVAL_SHAPE;  // the MN-major B stage holds 128 columns per CTA
  int rc = fa::check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pair) return a_from_tmem ? launch_mma_rate<true, true>(n, groups, ctas, st) : launch_mma_rate<true, false>(n, groups, ctas, st);
  return a_from_tmem ? launch_mma_rate<false, true>(n, groups, ctas, st) : launch_mma_rate<false, false>(n, groups, ctas, st);
}

extern "C" int fa_sm100_probe_reduce_rate(float* acc, int slices, int nqt, int nkt, int flags, void* stream) {
  if (slices <= 0 || nqt <= 0 || nkt <= 0) return FA_SM100_EINVAL_SHAPE;
  if (!fa::aligned16(acc)) return FA_SM100_EINVAL_PTR;
  int rc = fa::check_device();
  if (rc) return rc;
  CUtensorMap tm;
  const uint64_t rows = static_cast<uint64_t>(nqt) * 128;
  if ((rc = fa::make_tmap_3d(&tm, acc, fa::kElemF32, 128, rows, static_cast<uint64_t>(slices), rows * 128, 32, 128)))
    return rc;
  const int smem = (flags & 4) ? 200 * 1024 : 65536;  // flag 4: one CTA per SM, like the backward kernel
  cudaFuncSetAttribute(fa::fa_reduce_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  fa::fa_reduce_rate_kernel<<<slices * nkt, 128, smem, static_cast<cudaStream_t>(stream)>>>(tm, acc, nqt, nkt, flags);
  return fa::launch_status();
}
