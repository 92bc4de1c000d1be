// FP8 (e4m3) attention forward for sm_100a, head dim 128 -- the real path behind the reference's `fp8=True` flag
// (src/fa3/op.py:7, src/fa3/cuda/impl.py:40-55).  The reference only EMULATES fp8 (src/fa3/torch/impl.py:20-72,123-131:
// sign-flip + Hadamard "incoherent processing" of Q and K, per-block absmax scales, quantise -> dequantise, then the
// ordinary forward); here the same recipe feeds the tensor cores:
//   1. fa_fp8_quant_kernel: Q and K rows are sign-flipped (Philox bits from `seed`), Hadamard-transformed along the
//      head dim and divided by sqrt(d) (Q K^T is unchanged: H H^T = d I); Q, K and V are then scaled per 128-row block
//      by absmax / 448 and rounded to e4m3.  Scales go to fp32 side arrays (one per block).
//   2. fa_fwd_fp8_kernel: the main forward's structure (two 128-row query tiles ping-pong, TMA ring, S and O in TMEM)
//      with `tcgen05.mma kind::f8f6f4` for both products.  S = Q8 K8^T is rescaled by sq[i] * sk[j] inside the exp2
//      argument (the running row max is kept in scaled log2 units because the factor changes per K/V block);
//      P is multiplied by 64 * sv[j] / sv_ref (sv_ref = the slice's largest V scale) and rounded to e4m3 as the A
//      operand of the P V product, so V's per-block scale rides on P and O only needs sv_ref / 64 at the end.
// An e4m3 tile row is 128 bytes = exactly one 128-byte swizzle row, so a Q/K/V tile is ONE 16 KiB TMA box and the
// K-major products step their descriptors by the same 32 bytes per instruction as the 16-bit kernels.
#include "ptx.cuh"
#include "fa_host.cuh"

namespace fa {

constexpr int kF8D = 128, kF8BM = 128, kF8BN = 128, kF8Step = 64, kF8Threads = 352;
// P = 2^(s - m_ref) is stored in e4m3 as P * 64 * (sv[j] / sv_ref).  The lazy rescale lets the running reference lag the
// true row max by up to 2^2, so P <= 4 and the stored value <= 256 < 448 (e4m3 max); the smallest representable
// probability is 2^-9 / 64 = 2^-15 of the reference.  (The 16-bit kernels lag by up to 2^8: bf16 has the range, e4m3
// does not.)
constexpr float kF8PScale = 64.f;
constexpr float kF8RescaleThreshold = 2.f;

struct Fp8Params {
  float* lse;
  const float* sq;      // (bh, nqt)  per 128-row block scales of Q8, K8, V8
  const float* sk;      // (bh, nkt)
  const float* sv;      // (bh, nkt)
  const float* sv_ref;  // (bh)       max over the slice's V scales
  long long lse_bh_stride;
  int n_q, n_kv, bh, causal, diag, npairs, group_log2, nqt, nkt;
  float scale_log2;
};

struct Fp8Cfg {
  static constexpr int kStages = 8;
  static constexpr int kTileBytes = 128 * kF8D;       // one e4m3 Q / K / V tile: 16 KiB, one swizzled sub-tile
  static constexpr int kOutSub = 128 * 128;           // 64 output columns (16-bit) of one query tile
  static constexpr int kOutBytes = 2 * kOutSub;       // O staging of one query tile
  static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 2 * kOutBytes + 1024 + 256;
};
static_assert(Fp8Cfg::kSmemBytes <= 232448, "fp8 forward smem budget");

__device__ __forceinline__ int fp8_num_steps(int row0, const Fp8Params& p) {
  if (row0 >= p.n_q) return 0;
  int n = (p.n_kv + kF8Step - 1) / kF8Step;
  if (p.causal) {
    const long long last_visible = static_cast<long long>(row0) + kF8BM - 1 + p.diag;
    if (last_visible < 0) return 0;
    const int nc = static_cast<int>(last_visible / kF8Step) + 1;
    n = nc < n ? nc : n;
  }
  return n;
}

#ifndef FA_GRID_Y_BITS
#define FA_GRID_Y_BITS 15
#endif

template <bool kBF16>
__global__ void __launch_bounds__(kF8Threads, 1)
fa_fwd_fp8_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const Fp8Params p) {
  using Cfg = Fp8Cfg;
  constexpr int NS = Cfg::kStages, D = kF8D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                  // 2 tiles
  uint8_t* kv_smem = smem + 2 * Cfg::kTileBytes;           // NS tiles
  uint8_t* o_smem = kv_smem + NS * Cfg::kTileBytes;        // 2 x 32 KiB staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_smem + 2 * Cfg::kOutBytes);
  uint64_t* q_full = bars;             // [2]
  uint64_t* s_full = bars + 2;         // [2][2]
  uint64_t* p_ready = bars + 6;        // [2][2]
  uint64_t* pv_done = bars + 10;       // [2][2]
  uint64_t* kv_full = bars + 14;       // [NS]
  uint64_t* kv_empty = bars + 14 + NS; // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14 + 2 * NS);

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ((blockIdx.x >> p.group_log2) << FA_GRID_Y_BITS) + blockIdx.y;
  const int bh = static_cast<int>((blockIdx.z << p.group_log2) + (blockIdx.x & ((1u << p.group_log2) - 1u)));
  if (bh >= p.bh || rank >= static_cast<uint32_t>(p.npairs)) return;
  const int pair = p.npairs - 1 - static_cast<int>(rank);
  const int row0_t0 = pair * 2 * kF8BM;
  const int nt0 = fp8_num_steps(row0_t0, p);
  const int nt1 = fp8_num_steps(row0_t0 + kF8BM, p);
  const int ntmax = nt0 > nt1 ? nt0 : nt1;
  const int n_kv_tiles = (ntmax + 1) >> 1;

  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&q_full[i], 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int i = 0; i < 2; ++i) {
      mbar_arrive_expect_tx(&q_full[i], Cfg::kTileBytes);
      tma_load_3d(q_smem + i * Cfg::kTileBytes, &tm_q, &q_full[i], 0, row0_t0 + i * kF8BM, bh);
    }
    for (int t = 0; t < 2 * n_kv_tiles && t < NS; ++t) {
      mbar_arrive_expect_tx(&kv_full[t], Cfg::kTileBytes);
      tma_load_3d(kv_smem + t * Cfg::kTileBytes, (t & 1) ? &tm_v : &tm_k, &kv_full[t], 0, (t >> 1) * kF8BN, bh);
    }
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 8) {
    if (lane == 0) {
      for (int t = NS; t < 2 * n_kv_tiles; ++t) {
        const int stage = t % NS;
        mbar_wait(&kv_empty[stage], ((t / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::kTileBytes);
        tma_load_3d(kv_smem + stage * Cfg::kTileBytes, (t & 1) ? &tm_v : &tm_k, &kv_full[stage], 0, (t >> 1) * kF8BN, bh);
      }
    }
    __syncwarp();
  } else if (warp >= 9) {
    // ===================================== MMA issuers (warp 9 -> tile 0, warp 10 -> tile 1) =====================
    const int i = warp - 9;
    const int nti = i == 0 ? nt0 : nt1;
    if (ntmax > 0) {
      constexpr uint32_t idesc_s = umma_idesc(false, kF8BM, kF8Step, false, false);  // e4m3 x e4m3, both K-major, N = 64
      constexpr uint32_t idesc_o = umma_idesc(false, kF8BM, D, false, true);         // P (TMEM) x V MN-major, N = 128
      constexpr uint32_t kStageLo = Cfg::kTileBytes >> 4;
      constexpr uint32_t kHalfLo = (kF8Step * 128) >> 4;  // second 64 key rows of a tile
      const uint32_t q_lo = umma_desc_lo(smem_u32(q_smem) + i * Cfg::kTileBytes, 16);
      const uint32_t k_lo0 = umma_desc_lo(smem_u32(kv_smem), 16);
      const uint32_t v_lo0 = umma_desc_lo(smem_u32(kv_smem), Cfg::kTileBytes);
      const uint32_t t_s = tmem_base + i * 128;
      const uint32_t t_o = tmem_base + 256 + i * D;
      auto issue_s = [&](int step, uint32_t stage) {
        const uint32_t b_lo = k_lo0 + stage * kStageLo + (step & 1) * kHalfLo;
        const uint32_t d_tmem = t_s + (step & 1) * kF8Step;
#pragma unroll
        for (int kk = 0; kk < D / 32; ++kk)  // 32 bytes of the head dim per instruction
          umma_ss_f8(d_tmem, umma_desc(q_lo + kk * 2), umma_desc(b_lo + kk * 2), idesc_s, kk > 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int step, uint32_t stage, bool acc) {
        const uint32_t b_lo = v_lo0 + stage * kStageLo + (step & 1) * kHalfLo;
        const uint32_t a_tmem = t_s + (step & 1) * kF8Step;
#pragma unroll
        for (int kk = 0; kk < kF8Step / 32; ++kk)  // A: 32 keys = 8 TMEM columns; B: 32 key rows = 4 KiB
          umma_ts_f8(t_o, a_tmem + kk * 8, umma_desc(b_lo + kk * 256), idesc_o, (acc || kk > 0) ? 1u : 0u);
      };
      auto stage_of = [&](uint32_t slot) { return slot & (NS - 1); };
      auto phase_of = [&](uint32_t slot) { return (slot / NS) & 1u; };

      mbar_wait(&q_full[i], 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        if (0 < nti) {
          issue_s(0, 0);
          tc_commit(&s_full[i * 2]);
        }
        if (1 < nti) {
          issue_s(1, 0);
          tc_commit(&s_full[i * 2 + 1]);
        }
        tc_commit(&kv_empty[0]);
      }
      __syncwarp();
      for (int t = 0; t < ntmax; ++t) {
        const uint32_t sv = 2 * (t >> 1) + 1;
        const int s2 = t + 2;
        const uint32_t sk = 2 * (s2 >> 1);
        if ((t & 1) == 0) mbar_wait(&kv_full[stage_of(sv)], phase_of(sv));
        if ((s2 & 1) == 0 && s2 < ntmax) mbar_wait(&kv_full[stage_of(sk)], phase_of(sk));
        if (t < nti) mbar_wait(&p_ready[i * 2 + (t & 1)], (t >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (t < nti) {
            issue_pv(t, stage_of(sv), t > 0);
            tc_commit(&pv_done[i * 2 + (t & 1)]);
          }
          if (s2 < nti) {
            issue_s(s2, stage_of(sk));
            tc_commit(&s_full[i * 2 + (s2 & 1)]);
          }
          if ((t & 1) == 1 || t == ntmax - 1) tc_commit(&kv_empty[stage_of(sv)]);
          if (s2 <= ntmax - 1 && ((s2 & 1) == 1 || s2 == ntmax - 1)) tc_commit(&kv_empty[stage_of(sk)]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax warpgroups =====================================
    const int wg = warp >> 2;
    const int row = threadIdx.x & 127;
    const int nt = wg == 0 ? nt0 : nt1;
    const int tile_row0 = row0_t0 + wg * kF8BM;
    const int row_l = tile_row0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_sel + wg * 128;
    const uint32_t t_o = tmem_base + lane_sel + 256 + wg * D;
    int vis = p.n_kv - 1;
    if (p.causal) {
      const long long cv = static_cast<long long>(row_l) + p.diag;
      vis = cv < vis ? static_cast<int>(cv < -1 ? -1 : cv) : vis;
    }
    const int q_tile = tile_row0 / kF8BM;
    const float cq = (q_tile < p.nqt ? p.sq[static_cast<long long>(bh) * p.nqt + q_tile] : 0.f) * p.scale_log2;
    const float sv_ref = p.sv_ref[bh];
    const float* sk_b = p.sk + static_cast<long long>(bh) * p.nkt;
    const float* sv_b = p.sv + static_cast<long long>(bh) * p.nkt;
    const float inv_ref = sv_ref > 0.f ? kF8PScale / sv_ref : 0.f;

    float m_ref = -INFINITY;  // running row max in SCALED log2 units (the scale changes per K/V block)
    float l_sum = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int buf = j & 1;
      const uint32_t t_sb = t_s + buf * kF8Step;
      const float c = cq * sk_b[j >> 1];          // log2-domain factor of this K/V block (> 0)
      const float pmul = sv_b[j >> 1] * inv_ref;  // <= 64: V's block scale rides on P
      mbar_wait(&s_full[wg * 2 + buf], (j >> 1) & 1);
      tc_fence_after();
      float s[kF8Step];
      tmem_ld32(t_sb, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(t_sb + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tc_wait_ld();
      const int lim = vis - j * kF8Step;
      if (lim < kF8Step - 1) {
#pragma unroll
        for (int x = 0; x < kF8Step; ++x) s[x] = (x > lim) ? -INFINITY : s[x];
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int x = 4; x < kF8Step; x += 4) {
        mx0 = fmaxf(mx0, s[x]);
        mx1 = fmaxf(mx1, s[x + 1]);
        mx2 = fmaxf(mx2, s[x + 2]);
        mx3 = fmaxf(mx3, s[x + 3]);
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      const float m_new = fmaxf(m_ref, m_tile);
      bool rescale = false;
      float alpha = 1.f;
      if (j == 0) {
        m_ref = m_new;
      } else {
        const bool need = (m_new - m_ref) > kF8RescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = ex2(m_ref - m_safe);
          l_sum *= alpha;
          m_ref = m_new;
          rescale = true;
        }
      }
      const float mc = (m_ref == -INFINITY) ? 0.f : m_ref;
      float2 ls_a = make_float2(0.f, 0.f), ls_b = make_float2(0.f, 0.f);
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
      uint32_t pk[16];  // 64 e4m3 values
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float2 ta = ffma2(make_float2(s[4 * x], s[4 * x + 1]), c2, nmc2);
        const float2 tb = ffma2(make_float2(s[4 * x + 2], s[4 * x + 3]), c2, nmc2);
        float2 pa, pb;
        pa.x = ex2(ta.x);
        pa.y = ex2(ta.y);
        pb.x = ex2(tb.x);
        pb.y = ex2(tb.y);
        ls_a = fadd2(ls_a, pa);
        ls_b = fadd2(ls_b, pb);
        pk[x] = pack4_e4m3(pa.x * pmul, pa.y * pmul, pb.x * pmul, pb.y * pmul);
      }
      tmem_st16(t_sb, pk);
      l_sum += (ls_a.x + ls_a.y) + (ls_b.x + ls_b.y);
      if (rescale) {
        mbar_wait(&pv_done[wg * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int q4 = 0; q4 < D / 32; ++q4) {
          float o[32];
          tmem_ld32(t_o + q4 * 32, reinterpret_cast<uint32_t*>(o));
          tc_wait_ld();
#pragma unroll
          for (int x = 0; x < 32; ++x) o[x] *= alpha;
          tmem_st32(t_o + q4 * 32, reinterpret_cast<const uint32_t*>(o));
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[wg * 2 + buf]);
    }
    // ------------------------------- epilogue -------------------------------
    if (nt > 0) {
      if (nt > 1) mbar_wait(&pv_done[wg * 2 + ((nt - 2) & 1)], ((nt - 2) >> 1) & 1);
      mbar_wait(&pv_done[wg * 2 + ((nt - 1) & 1)], ((nt - 1) >> 1) & 1);
      tc_fence_after();
    }
    const bool has_mass = l_sum > 0.f;
    const float w = has_mass ? sv_ref / (kF8PScale * l_sum) : 0.f;  // undo P's 64 / sv_ref, then normalise
    const float m_fin = (m_ref == -INFINITY) ? 0.f : m_ref;
    const float lse_val = has_mass ? (m_fin + log2f(l_sum)) * 0.6931471805599453f : -INFINITY;
    uint8_t* stage_tile = o_smem + wg * Cfg::kOutBytes;
#pragma unroll
    for (int q4 = 0; q4 < D / 32; ++q4) {
      float o[32];
      if (nt > 0) {
        tmem_ld32(t_o + q4 * 32, reinterpret_cast<uint32_t*>(o));
        tc_wait_ld();
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x) o[x] = 0.f;
      }
      uint32_t pk[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) pk[x] = pack2<kBF16>(o[2 * x] * w, o[2 * x + 1] * w);
      uint8_t* sub = stage_tile + (q4 >> 1) * Cfg::kOutSub + row * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int chunk = (q4 & 1) * 4 + ch;
        *reinterpret_cast<uint4*>(sub + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
      }
    }
    if (row_l < p.n_q) p.lse[static_cast<long long>(bh) * p.lse_bh_stride + row_l] = lse_val;
    fence_proxy_async_smem();
    named_bar_sync(1 + wg, 128);
    if (row == 0 && tile_row0 < p.n_q) {
      for (int ch = 0; ch < 2; ++ch) tma_store_3d(&tm_o, stage_tile + ch * Cfg::kOutSub, ch * 64, tile_row0, bh);
      tma_store_commit();
      tma_store_wait_exit();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Quantisation pre-pass (HBM-bound): one CTA of 512 threads per (128-row block, slice); warp w handles rows
// 8 w .. 8 w + 7, lane l holds head dim elements 4 l .. 4 l + 3.  kHadamard: x <- H (s o x) / sqrt(128) with s = +-1 from
// Philox (reference src/fa3/torch/impl.py:41-59: sign flip, Walsh-Hadamard butterflies, 1/sqrt(d); here a true WHT).
// Then scale = absmax / 448 over the block (reference :20-31), x / scale rounded to e4m3 (saturating).
// Two passes over the rows: the first computes the block maximum, the second RECOMPUTES the transform (the 32 KiB
// block comes back from L1/L2) and quantises -- no shared-memory copy of the block, so many CTAs fit per SM.
// ------------------------------------------------------------------------------------------------
template <bool kBF16, bool kHadamard>
__device__ __forceinline__ void fp8_load_transform(const uint16_t* __restrict__ row_ptr, int lane, const float (&sign)[4],
                                                   float (&v)[4]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(row_ptr + lane * 4);
  const float2 a = unpack2<kBF16>(raw.x), b = unpack2<kBF16>(raw.y);
  v[0] = a.x * sign[0];
  v[1] = a.y * sign[1];
  v[2] = b.x * sign[2];
  v[3] = b.y * sign[3];
  if (kHadamard) {
    // strides 1 and 2 inside the lane, 4 .. 64 across lanes
    const float t0 = v[0] + v[1], t1 = v[0] - v[1], t2 = v[2] + v[3], t3 = v[2] - v[3];
    v[0] = t0 + t2;
    v[1] = t1 + t3;
    v[2] = t0 - t2;
    v[3] = t1 - t3;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float other = __shfl_xor_sync(0xffffffffu, v[i], m);
        v[i] = (lane & m) ? other - v[i] : v[i] + other;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= 0.08838834764831845f;  // 1 / sqrt(128)
  }
}

template <bool kBF16, bool kHadamard>
__global__ void __launch_bounds__(512) fa_fp8_quant_kernel(const uint16_t* __restrict__ x, uint8_t* __restrict__ out,
                                                           float* __restrict__ scales, int n, long long x_bh_stride,
                                                           int n_tiles, uint32_t seed_lo, uint32_t seed_hi) {
  __shared__ float warp_max[16];
  const int tile = blockIdx.x, bh = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float sign[4] = {1.f, 1.f, 1.f, 1.f};
  if (kHadamard) {
    const Philox4 bits = philox4x32_7(static_cast<uint32_t>(lane), 0u, 0u, 0u, seed_lo, seed_hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) sign[i] = (bits.w[i] & 1u) ? -1.f : 1.f;
  }
  const uint16_t* xb = x + static_cast<long long>(bh) * x_bh_stride;
  const int row0 = tile * 128 + warp * 8;
  float amax = 0.f;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    if (row0 + rr < n) {
      float v[4];
      fp8_load_transform<kBF16, kHadamard>(xb + static_cast<long long>(row0 + rr) * 128, lane, sign, v);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
  if (lane == 0) warp_max[warp] = amax;
  __syncthreads();
  amax = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) amax = fmaxf(amax, warp_max[w]);
  const float scale = amax > 0.f ? amax / 448.f : 1.f;
  const float inv = 1.f / scale;
  if (threadIdx.x == 0) scales[static_cast<long long>(bh) * n_tiles + tile] = scale;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    if (row0 + rr < n) {
      float v[4];
      fp8_load_transform<kBF16, kHadamard>(xb + static_cast<long long>(row0 + rr) * 128, lane, sign, v);
      *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(bh) * n + row0 + rr) * 128 + lane * 4) =
          pack4_e4m3(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
    }
  }
}

}  // namespace fa

extern "C" int fa_sm100_fp8_quantize(const void* x, void* out8, float* scales, int64_t bh, int64_t n, int32_t d,
                                     int64_t x_bh_stride, int32_t dtype, int32_t hadamard, uint64_t seed, void* stream) {
  if (dtype != FA_SM100_DTYPE_F16 && dtype != FA_SM100_DTYPE_BF16) return FA_SM100_EINVAL_DTYPE;
  if (d != 128) return FA_SM100_EINVAL_HEADDIM;
  if (bh <= 0 || n <= 0 || bh > 65535 || n > (1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  if (x_bh_stride == 0) x_bh_stride = n * d;
  if (x_bh_stride < n * d || (x_bh_stride % 8)) return FA_SM100_EINVAL_SHAPE;
  if (!fa::aligned16(x) || !fa::aligned16(out8) || scales == nullptr) return FA_SM100_EINVAL_PTR;
  int rc = fa::check_device();
  if (rc) return rc;
  const int n_tiles = static_cast<int>((n + 127) / 128);
  const dim3 grid(static_cast<unsigned>(n_tiles), static_cast<unsigned>(bh));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint16_t* xp = static_cast<const uint16_t*>(x);
  uint8_t* op = static_cast<uint8_t*>(out8);
  const uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
  const int nn = static_cast<int>(n);
  const bool bf = dtype == FA_SM100_DTYPE_BF16;
#define FA_QUANT_GO(BF, HAD) \
  fa::fa_fp8_quant_kernel<BF, HAD><<<grid, 512, 0, st>>>(xp, op, scales, nn, x_bh_stride, n_tiles, lo, hi)
  if (hadamard) {
    if (bf) FA_QUANT_GO(true, true); else FA_QUANT_GO(false, true);
  } else {
    if (bf) FA_QUANT_GO(true, false); else FA_QUANT_GO(false, false);
  }
#undef FA_QUANT_GO
  return fa::launch_status();
}

extern "C" int fa_sm100_fwd_fp8(const fa_sm100_shape* s, const void* q8, const void* k8, const void* v8,
                                const float* sq, const float* sk, const float* sv, const float* sv_ref, void* o,
                                float* lse, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g);
  if (rc) return rc;
  if (g.d != 128) return FA_SM100_EINVAL_HEADDIM;
  if (!fa::aligned16(q8) || !fa::aligned16(k8) || !fa::aligned16(v8) || !fa::aligned16(o) || lse == nullptr ||
      sq == nullptr || sk == nullptr || sv == nullptr || sv_ref == nullptr)
    return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  using Cfg = fa::Fp8Cfg;
  CUtensorMap tm_q, tm_k, tm_v, tm_o;
  // the e4m3 tensors are dense (bh, n, 128) bytes
  if ((rc = fa::make_tmap_3d(&tm_q, q8, fa::kElemU8, 128, g.n_q, g.bh, g.n_q * 128, 128, 128))) return rc;
  if ((rc = fa::make_tmap_3d(&tm_k, k8, fa::kElemU8, 128, g.n_kv, g.bh, g.n_kv * 128, 128, 128))) return rc;
  if ((rc = fa::make_tmap_3d(&tm_v, v8, fa::kElemU8, 128, g.n_kv, g.bh, g.n_kv * 128, 128, 128))) return rc;
  const bool bf = g.dtype == FA_SM100_DTYPE_BF16;
  if ((rc = fa::make_tmap_3d(&tm_o, o, bf ? fa::kElemBF16 : fa::kElemF16, 128, g.n_q, g.bh, g.q_bh_stride, 64, 128)))
    return rc;
  fa::Fp8Params p;
  p.lse = lse;
  p.sq = sq;
  p.sk = sk;
  p.sv = sv;
  p.sv_ref = sv_ref;
  p.lse_bh_stride = g.lse_bh_stride;
  p.n_q = static_cast<int>(g.n_q);
  p.n_kv = static_cast<int>(g.n_kv);
  p.bh = static_cast<int>(g.bh);
  p.causal = g.causal;
  p.diag = g.diag;
  p.nqt = static_cast<int>((g.n_q + 127) / 128);
  p.nkt = static_cast<int>((g.n_kv + 127) / 128);
  p.npairs = static_cast<int>((g.n_q + 255) / 256);
  p.group_log2 = fa::sched_group_log2(g.causal != 0, p.npairs, g.bh);
  while (((g.bh + (1ll << p.group_log2) - 1) >> p.group_log2) > 65535) ++p.group_log2;
  p.scale_log2 = g.scale * 1.4426950408889634f;
  const long long rank_lo = p.npairs < (1 << FA_GRID_Y_BITS) ? p.npairs : (1 << FA_GRID_Y_BITS);
  const long long rank_hi = (p.npairs + (1 << FA_GRID_Y_BITS) - 1) >> FA_GRID_Y_BITS;
  const long long gx = rank_hi << p.group_log2, gz = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  if (gx > 0x7fffffffll || gz > 65535) return FA_SM100_EINVAL_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rank_lo), static_cast<unsigned>(gz));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bf) {
    if (cudaFuncSetAttribute(fa::fa_fwd_fp8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    fa::fa_fwd_fp8_kernel<true><<<grid, fa::kF8Threads, Cfg::kSmemBytes, st>>>(tm_q, tm_k, tm_v, tm_o, p);
  } else {
    if (cudaFuncSetAttribute(fa::fa_fwd_fp8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    fa::fa_fwd_fp8_kernel<false><<<grid, fa::kF8Threads, Cfg::kSmemBytes, st>>>(tm_q, tm_k, tm_v, tm_o, p);
  }
  return fa::launch_status();
}
