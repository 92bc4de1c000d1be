// Attention backward for head dims 129..256 on sm_100a (the reference's benchmark sweeps
// --head-dim 64 128 256 forward AND backward, benchmarks/bench_utils.py:256).
//
// dK and dV accumulators of 256 columns each would fill TMEM, so the main backward's layout (kv rows on TMEM lanes)
// does not extend.  Here the CTA owns a 64-row K/V tile and keeps dK^T and dV^T -- head dim on the lanes, two 128-lane
// halves of 64 columns each = 256 columns -- and walks the visible 128-row query tiles with the QUERY rows on the
// lanes for everything else:
//   S    = Q  K^T            A = Q  smem K-major,       B = K smem K-major, N = 64      -> TMEM [0,64)
//   dP   = dO V^T            A = dO smem K-major,       B = V smem K-major, N = 64      -> TMEM [64,128)
//   P = 2^(S c - lse log2e), dS = P o (dP - delta): thread = query row, so lse / delta are per-thread scalars;
//     P, dS -> 16-bit shared-memory tiles [q][kv]; dS also packed over the consumed dP columns (A operand of dQ)
//   dV^T += dO^T P           A = dO smem read MN-major (M = head dim, two halves), B = P  smem MN-major, N = 64
//   dK^T += Q^T  dS          A = Q  smem read MN-major,                           B = dS smem MN-major, N = 64
//   dQ_h  = dS K[:, h]       A = dS in TMEM, B = K smem MN-major, N = 128, one half of the head dim at a time
//                            -> TMEM [128,256), drained to the fp32 dq_accum by TMA reduce-add (staged in the P / dS
//                            tiles, which are dead by then)
// TMEM: S 64 | dP 64 | dQ half 128 | dV^T 2 x 64 | dK^T 2 x 64 = 512 columns.
// SMEM: K 32 + V 32 + Q 64 + dO 64 + P 16 + dS 16 KiB = 224 KiB: Q and dO are single-buffered (there is no room for a
// second stage at this head dim), so loads, products and the softmax phase overlap only across the dQ drains.  This
// is a capability kernel (roughly 0.2 of the tensor peak), not a tuned one.
// Warps: 0-3 softmax + dQ drain + epilogue (thread = query row / head-dim lane), 4 TMA producer, 5 MMA issuer.
#include "ptx.cuh"
#include "fa_host.cuh"

namespace fa {

constexpr int kB2Threads = 192;
constexpr int kB2Q = 128;   // query rows per tile
constexpr int kB2KV = 64;   // key rows per tile (= CTA)
constexpr int kB2D = 256;

struct Bwd256Params {
  const float* rowstats;  // (bh, nqt, 2, 128) as written by fa_sm100_bwd_prepare
  uint16_t* dk;
  uint16_t* dv;
  long long kv_bh_stride;
  int n_q, n_kv, bh, causal, diag, nqt, nkt, group_log2, d;
  float scale_log2, scale;
};

struct Bwd256Cfg {
  static constexpr int kQSub = kB2Q * 128;             // 16 KiB: 64 head-dim columns of a query tile
  static constexpr int kKVSub = kB2KV * 128;           // 8 KiB
  static constexpr int kQBytes = 4 * kQSub;            // 64 KiB
  static constexpr int kKVBytes = 4 * kKVSub;          // 32 KiB
  static constexpr int kPBytes = kB2Q * 128;           // P / dS tile: 128 query rows x 64 keys, 16-bit = 16 KiB
  static constexpr int kOffK = 0, kOffV = kKVBytes, kOffQ = 2 * kKVBytes, kOffDO = kOffQ + kQBytes;
  static constexpr int kOffP = kOffDO + kQBytes, kOffDS = kOffP + kPBytes, kOffStats = kOffDS + kPBytes;
  static constexpr int kOffBars = kOffStats + 1024;
  static constexpr int kSmemBytes = kOffBars + 256;
};
static_assert(Bwd256Cfg::kSmemBytes <= 232448, "d=256 backward smem budget");

enum B2Bar : int { kB2KVFull = 0, kB2QFull, kB2DOFull, kB2SFull, kB2PdsReady, kB2GradsDone, kB2DQFull, kB2DQDrained,
                   kB2AllDone, kB2Count };

#ifndef FA_GRID_Y_BITS
#define FA_GRID_Y_BITS 15
#endif

template <bool kBF16>
__global__ void __launch_bounds__(kB2Threads, 1)
fa_bwd256_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_dq, const Bwd256Params p) {
  using Cfg = Bwd256Cfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* k_smem = smem + Cfg::kOffK;
  uint8_t* v_smem = smem + Cfg::kOffV;
  uint8_t* q_smem = smem + Cfg::kOffQ;
  uint8_t* do_smem = smem + Cfg::kOffDO;
  uint8_t* p_smem = smem + Cfg::kOffP;
  uint8_t* ds_smem = smem + Cfg::kOffDS;
  float* stats_smem = reinterpret_cast<float*>(smem + Cfg::kOffStats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kB2Count);

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int j = static_cast<int>(((blockIdx.x >> p.group_log2) << FA_GRID_Y_BITS) + blockIdx.y);  // 64-row kv tile
  const int bh = static_cast<int>((blockIdx.z << p.group_log2) + (blockIdx.x & ((1u << p.group_log2) - 1u)));
  if (bh >= p.bh || j >= p.nkt) return;
  int i_min = 0;
  if (p.causal) {
    const int first = j * kB2KV - p.diag;  // first query row that sees this tile's first key
    i_min = first > 0 ? first / kB2Q : 0;
  }
  const int n_iter = p.nqt > i_min ? p.nqt - i_min : 0;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("fa_sm100 bwd256: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  auto load_q = [&](int i) {
    mbar_arrive_expect_tx(&bars[kB2QFull], Cfg::kQBytes + 1024);
    for (int c = 0; c < 4; ++c) tma_load_3d(q_smem + c * Cfg::kQSub, &tm_q, &bars[kB2QFull], c * 64, i * kB2Q, bh);
    bulk_load_1d(stats_smem, p.rowstats + (static_cast<long long>(bh) * p.nqt + i) * 256, 1024, &bars[kB2QFull]);
  };
  auto load_do = [&](int i) {
    mbar_arrive_expect_tx(&bars[kB2DOFull], Cfg::kQBytes);
    for (int c = 0; c < 4; ++c) tma_load_3d(do_smem + c * Cfg::kQSub, &tm_do, &bars[kB2DOFull], c * 64, i * kB2Q, bh);
  };
  if (warp == 4 && lane == 0) {
    for (int b = 0; b < kB2Count; ++b) mbar_init(&bars[b], (b == kB2PdsReady || b == kB2DQDrained) ? 128u : 1u);
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_arrive_expect_tx(&bars[kB2KVFull], 2 * Cfg::kKVBytes);
    for (int c = 0; c < 4; ++c) {
      tma_load_3d(k_smem + c * Cfg::kKVSub, &tm_k, &bars[kB2KVFull], c * 64, j * kB2KV, bh);
      tma_load_3d(v_smem + c * Cfg::kKVSub, &tm_v, &bars[kB2KVFull], c * 64, j * kB2KV, bh);
    }
    if (n_iter > 0) {
      load_q(i_min);
      load_do(i_min);
    }
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  constexpr uint32_t kColS = 0, kColDP = 64, kColDQ = 128, kColDVT = 256, kColDKT = 384;

  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      for (int it = 1; it < n_iter; ++it) {
        mbar_wait(&bars[kB2GradsDone], (it - 1) & 1);  // dV^T / dK^T of the previous tile have read Q and dO
        load_q(i_min + it);
        load_do(i_min + it);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================================== MMA issuer =====================================
    if (n_iter > 0) {
      constexpr uint32_t idesc_s = umma_idesc(kBF16, kB2Q, kB2KV, false, false);   // [q x kv], both K-major
      constexpr uint32_t idesc_t = umma_idesc(kBF16, 128, kB2KV, true, true);       // [d-half x kv], A and B MN-major
      constexpr uint32_t idesc_q = umma_idesc(kBF16, kB2Q, 128, false, true);       // [q x d-half], A in TMEM, B MN-major
      const uint32_t q_km = umma_desc_lo(smem_u32(q_smem), 16), do_km = umma_desc_lo(smem_u32(do_smem), 16);
      const uint32_t k_km = umma_desc_lo(smem_u32(k_smem), 16), v_km = umma_desc_lo(smem_u32(v_smem), 16);
      const uint32_t q_mn = umma_desc_lo(smem_u32(q_smem), Cfg::kQSub), do_mn = umma_desc_lo(smem_u32(do_smem), Cfg::kQSub);
      const uint32_t k_mn = umma_desc_lo(smem_u32(k_smem), Cfg::kKVSub);
      const uint32_t p_mn = umma_desc_lo(smem_u32(p_smem), Cfg::kPBytes), ds_mn = umma_desc_lo(smem_u32(ds_smem), Cfg::kPBytes);
      // D[q, kv] = A[q, :] . B[kv, :]  over the 256 head-dim columns (16 instructions of 16)
      auto mma_scores = [&](uint32_t d_col, uint32_t a_lo, uint32_t b_lo) {
#pragma unroll
        for (int kk = 0; kk < kB2D / 16; ++kk)
          umma_ss(tmem_base + d_col, umma_desc(a_lo + (kk >> 2) * (Cfg::kQSub >> 4) + (kk & 3) * 2),
                  umma_desc(b_lo + (kk >> 2) * (Cfg::kKVSub >> 4) + (kk & 3) * 2), idesc_s, kk > 0 ? 1u : 0u);
      };
      // D[d-half h, kv] (+)= A^T . B: A = the query-major tile read MN-major (two 64-column sub-tiles of half h),
      // B = the [q][kv] tile read MN-major; contraction over the 128 query rows (8 instructions of 16 rows = 2 KiB)
      auto mma_transposed = [&](uint32_t d_col, uint32_t a_lo, uint32_t b_lo, int h, bool acc) {
#pragma unroll
        for (int kk = 0; kk < kB2Q / 16; ++kk)
          umma_ss(tmem_base + d_col + h * kB2KV, umma_desc(a_lo + h * ((2 * Cfg::kQSub) >> 4) + kk * 128),
                  umma_desc(b_lo + kk * 128), idesc_t, (acc || kk > 0) ? 1u : 0u);
      };
      // dQ[q, d-half h] = dS[q, kv] (TMEM, 16-bit, columns [64,96)) . K[kv, d-half h] (MN-major, two 8 KiB sub-tiles)
      auto mma_dq = [&](int h) {
#pragma unroll
        for (int kk = 0; kk < kB2KV / 16; ++kk)
          umma_ts(tmem_base + kColDQ, tmem_base + kColDP + kk * 8,
                  umma_desc(k_mn + h * ((2 * Cfg::kKVSub) >> 4) + kk * 128), idesc_q, kk > 0 ? 1u : 0u);
      };
      mbar_wait(&bars[kB2KVFull], 0);
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&bars[kB2QFull], it & 1);
        mbar_wait(&bars[kB2DOFull], it & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_scores(kColS, q_km, k_km);
          mma_scores(kColDP, do_km, v_km);
          tc_commit(&bars[kB2SFull]);
        }
        __syncwarp();
        mbar_wait(&bars[kB2PdsReady], it & 1);
        tc_fence_after();
        if (elect_one()) {
          for (int h = 0; h < 2; ++h) {
            mma_transposed(kColDVT, do_mn, p_mn, h, it > 0);   // dV^T += dO^T P
            mma_transposed(kColDKT, q_mn, ds_mn, h, it > 0);   // dK^T += Q^T dS
          }
          tc_commit(&bars[kB2GradsDone]);  // Q, dO (and the P / dS tiles) have been read
          mma_dq(0);
          tc_commit(&bars[kB2DQFull]);
        }
        __syncwarp();
        mbar_wait(&bars[kB2DQDrained], 0);  // two phases per tile: half 0 ...
        tc_fence_after();
        if (elect_one()) {
          mma_dq(1);
          tc_commit(&bars[kB2DQFull]);
        }
        __syncwarp();
        // ... half 1: the next tile's S / dP overwrite the packed dS operand, so dQ(1) must have COMPLETED (its own
        // commit), but its drain may still be running: S / dP do not touch the dQ columns, and the next dQ(0) is only
        // issued after pds_ready, which the draining warpgroup signals when it is done
        mbar_wait(&bars[kB2DQFull], 1);
      }
      tc_commit_elect(&bars[kB2AllDone]);
    }
  } else {
    // ===================================== softmax / drain / epilogue warpgroup =====================================
    const int r = threadIdx.x;  // query row inside the tile == TMEM lane (and head-dim lane in the epilogue)
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    for (int it = 0; it < n_iter; ++it) {
      const int i = i_min + it;
      mbar_wait(&bars[kB2QFull], it & 1);  // row statistics ride on the Q barrier
      mbar_wait(&bars[kB2SFull], it & 1);
      tc_fence_after();
      const float nl2 = stats_smem[r];         // -lse * log2e (-inf for rows past the end or without keys)
      const float nd = stats_smem[128 + r];    // -delta
      const int q_glob = i * kB2Q + r;
      // key (j*64 + c) is visible to this query iff c <= lim
      int lim = p.n_kv - 1 - j * kB2KV;
      if (p.causal) {
        const int cl = q_glob + p.diag - j * kB2KV;
        lim = cl < lim ? cl : lim;
      }
      float pr[64];
      {
        float s[64];
        tmem_ld32(tmem_base + lane_sel + kColS, reinterpret_cast<uint32_t*>(s));
        tmem_ld32(tmem_base + lane_sel + kColS + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tc_wait_ld();
#pragma unroll
        for (int x = 0; x < 64; ++x) {
          const float e = ex2(fmaf(s[x], p.scale_log2, nl2));
          pr[x] = (x <= lim) ? e : 0.f;
        }
      }
      uint32_t dsw[32];
      {
        float dp[64];
        tmem_ld32(tmem_base + lane_sel + kColDP, reinterpret_cast<uint32_t*>(dp));
        tmem_ld32(tmem_base + lane_sel + kColDP + 32, reinterpret_cast<uint32_t*>(dp) + 32);
        tc_wait_ld();
#pragma unroll
        for (int x = 0; x < 32; ++x)
          dsw[x] = pack2<kBF16>(pr[2 * x] * (dp[2 * x] + nd), pr[2 * x + 1] * (dp[2 * x + 1] + nd));
      }
      // P and dS tiles: row r = 128 bytes = 64 keys, 128-byte swizzle (chunk ^ (r & 7))
      uint8_t* prow = p_smem + r * 128;
      uint8_t* drow = ds_smem + r * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t pw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pw[e] = pack2<kBF16>(pr[ch * 8 + 2 * e], pr[ch * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        *reinterpret_cast<uint4*>(drow + ((ch ^ (r & 7)) << 4)) =
            make_uint4(dsw[4 * ch], dsw[4 * ch + 1], dsw[4 * ch + 2], dsw[4 * ch + 3]);
      }
      tmem_st32(tmem_base + lane_sel + kColDP, dsw);  // packed dS over the dP columns this thread has consumed
      tc_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bars[kB2PdsReady]);

      // drain the two halves of dQ: TMEM -> registers -> swizzled fp32 staging (the P / dS tiles: dead once dq_full
      // fires, because that commit also covers dV^T and dK^T) -> TMA reduce-add, 32 columns per box
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&bars[kB2DQFull], h);
        tc_fence_after();
#pragma unroll 1
        for (int half64 = 0; half64 < 2; ++half64) {
          float v[64];
          tmem_ld32(tmem_base + lane_sel + kColDQ + half64 * 64, reinterpret_cast<uint32_t*>(v));
          tmem_ld32(tmem_base + lane_sel + kColDQ + half64 * 64 + 32, reinterpret_cast<uint32_t*>(v) + 32);
          tc_wait_ld();
          if (half64 == 1) {  // last read of this dQ half: the MMA warp may overwrite the columns
            tc_fence_before();
            mbar_arrive(&bars[kB2DQDrained]);
          }
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            uint8_t* rowp = (cb == 0 ? p_smem : ds_smem) + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(rowp + ((c ^ (r & 7)) << 4)) =
                  make_float4(v[cb * 32 + 4 * c], v[cb * 32 + 4 * c + 1], v[cb * 32 + 4 * c + 2], v[cb * 32 + 4 * c + 3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (r == 0) {
            for (int cb = 0; cb < 2; ++cb) {
              const int col = h * 128 + half64 * 64 + cb * 32;
              if (col < p.d) tma_reduce_add_3d(&tm_dq, cb == 0 ? p_smem : ds_smem, col, i * kB2Q, bh);
            }
            tma_store_commit();
            tma_store_wait_read<0>();
          }
          named_bar_sync(1, 128);  // staging is free again (next 64 columns / next tile's P and dS)
        }
      }
    }
    // ------------------------------- epilogue: dK^T, dV^T (head dim on the lanes) -> dk, dv -------------------------------
    if (n_iter > 0) {
      mbar_wait(&bars[kB2AllDone], 0);
      tc_fence_after();
    }
    const int kv0 = j * kB2KV;
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {  // dV^T half 0, half 1, dK^T half 0, half 1
      const int h = m & 1;
      const bool is_k = m >= 2;
      const int dcol = h * 128 + r;
      float a[64];
      if (n_iter > 0) {
        const uint32_t t = tmem_base + lane_sel + (is_k ? kColDKT : kColDVT) + h * kB2KV;
        tmem_ld32(t, reinterpret_cast<uint32_t*>(a));
        tmem_ld32(t + 32, reinterpret_cast<uint32_t*>(a) + 32);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int x = 0; x < 64; ++x) a[x] = 0.f;
      }
      if (dcol < p.d) {
        uint16_t* dst = (is_k ? p.dk : p.dv) + static_cast<long long>(bh) * p.kv_bh_stride + dcol;
        const float mul = is_k ? p.scale : 1.f;
#pragma unroll
        for (int x = 0; x < 64; ++x) {
          if (kv0 + x < p.n_kv)
            dst[static_cast<long long>(kv0 + x) * p.d] = static_cast<uint16_t>(pack2<kBF16>(a[x] * mul, 0.f) & 0xFFFFu);
        }
      }
    }
    if (r == 0) tma_store_wait_exit();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

template <bool kBF16>
static int launch_bwd256(const Geometry& g, const void* q, const void* k, const void* v, const void* d_o,
                         const float* rowstats, float* dq_accum, void* dk, void* dv, cudaStream_t stream) {
  using Cfg = Bwd256Cfg;
  const int elem = kBF16 ? kElemBF16 : kElemF16;
  CUtensorMap tm_q, tm_k, tm_v, tm_do, tm_dq;
  int rc;
  if ((rc = make_tmap_3d(&tm_q, q, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kB2Q))) return rc;
  if ((rc = make_tmap_3d(&tm_do, d_o, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kB2Q))) return rc;
  if ((rc = make_tmap_3d(&tm_k, k, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kB2KV))) return rc;
  if ((rc = make_tmap_3d(&tm_v, v, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kB2KV))) return rc;
  if ((rc = make_tmap_3d(&tm_dq, dq_accum, kElemF32, g.d, g.n_q, g.bh, g.q_bh_stride, 32, kB2Q))) return rc;
  Bwd256Params p;
  p.rowstats = rowstats;
  p.dk = static_cast<uint16_t*>(dk);
  p.dv = static_cast<uint16_t*>(dv);
  p.kv_bh_stride = g.kv_bh_stride;
  p.n_q = static_cast<int>(g.n_q);
  p.n_kv = static_cast<int>(g.n_kv);
  p.bh = static_cast<int>(g.bh);
  p.causal = g.causal;
  p.diag = g.diag;
  p.d = g.d;
  p.nqt = static_cast<int>((g.n_q + kB2Q - 1) / kB2Q);
  p.nkt = static_cast<int>((g.n_kv + kB2KV - 1) / kB2KV);
  p.group_log2 = sched_group_log2(g.causal != 0, p.nkt, g.bh);
  while (((g.bh + (1ll << p.group_log2) - 1) >> p.group_log2) > 65535) ++p.group_log2;
  p.scale = g.scale;
  p.scale_log2 = g.scale * 1.4426950408889634f;
  auto kern = fa_bwd256_kernel<kBF16>;
  static bool attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long rank_lo = p.nkt < (1 << FA_GRID_Y_BITS) ? p.nkt : (1 << FA_GRID_Y_BITS);
  const long long rank_hi = (p.nkt + (1 << FA_GRID_Y_BITS) - 1) >> FA_GRID_Y_BITS;
  const long long gx = rank_hi << p.group_log2, gz = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  if (gx > 0x7fffffffll || gz > 65535) return FA_SM100_EINVAL_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rank_lo), static_cast<unsigned>(gz));
  kern<<<grid, kB2Threads, Cfg::kSmemBytes, stream>>>(tm_q, tm_k, tm_v, tm_do, tm_dq, p);
  return launch_status();
}

// entry used by fa_sm100_bwd for 128 < d <= 256 (defined here, declared in fa_bwd_sm100.cu)
int bwd256_dispatch(const Geometry& g, const void* q, const void* k, const void* v, const void* d_o,
                    const float* rowstats, float* dq_accum, void* dk, void* dv, cudaStream_t st) {
  return g.dtype == FA_SM100_DTYPE_BF16 ? launch_bwd256<true>(g, q, k, v, d_o, rowstats, dq_accum, dk, dv, st)
                                        : launch_bwd256<false>(g, q, k, v, d_o, rowstats, dq_accum, dk, dv, st);
}

}  // namespace fa
