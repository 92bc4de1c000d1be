// Fused attention forward for sm_100a: TMA -> SMEM ring -> tcgen05.mma (S and O accumulators in TMEM) ->
// register softmax (one thread per query row) -> P back to TMEM as the A operand of the P*V MMA.
//
// Replaces the reference's host-side ATen tile loops fa{1,2,3}_forward (csrc/fa1/fa1_fwd.cu:30-107; algorithmic
// twin src/fa1/torch/impl.py:26-68): per query tile, stream KV tiles, online softmax with running max m and sum l,
// O = sum_j e^{S_j - m} V_j, finally O / l and lse = m + log l.
//
// CTA = two 128-row query tiles (ping-pong) of one (batch*head) slice.
//   warps 0-3  : softmax warpgroup for query tile 0 (thread t owns row t: TMEM lane == row, no shuffles)
//   warps 4-7  : softmax warpgroup for query tile 1
//   warp  8    : TMA producer (Q once, then K_0 V_0 K_1 V_1 ... through an NS-deep ring)
//   warps 9,10 : MMA issuers, one per query tile (independent issue streams); warp 9 owns the TMEM allocation
// TMEM (512 columns x 128 lanes, fp32):  [0,128) S0   [128,256) S1   [256,256+D) O0   [256+D,256+2D) O1
// The softmax works in STEPS of 64 key columns; each query tile's 128 S columns are two 64-column buffers, so
// S_i(t+1) is already in TMEM while the warpgroup is still exponentiating S_i(t) and never waits for the tensor pipe:
//   step t (KV tile t/2, half t%2, buffer t%2):  S_i(t) = Q_i K[half]^T  ->  P_i(t) (16-bit, first 32 columns of the
//   buffer, consumed straight from TMEM as the A operand)  ->  O_i += P_i(t) V[half]  ->  S_i(t+2) into the same buffer.
// The tensor pipe executes MMAs in issue order, so "P_i(t) V ; Q_i K(t+2)^T" needs no barrier in between even though
// S_i(t+2) overwrites P_i(t).
#include "ptx.cuh"
#include "fa_host.cuh"

namespace fa {

struct FwdParams {
  float* lse;
  const void* o_prev;
  const float* lse_prev;
  long long lse_bh_stride;
  long long o_bh_stride;  // elements; o_prev shares o's geometry
  long long n_items;      // work items = slice groups x tile pairs x slices per group (padding slices included)
  int n_q, n_kv, bh, causal, diag, npairs, group_log2;
  int d;             // true head dim (<= D): row stride of o_prev; columns [d, D) are zero-filled / clipped by TMA
  float scale_log2;  // softmax_scale * log2(e)
};

constexpr int kBM = 128;  // query rows per tile
constexpr int kBN = 128;  // key rows per K/V tile (TMA granularity)
constexpr int kStep = 64;  // key columns per softmax step
constexpr int kFwdThreads = 352;  // 8 softmax warps + producer + one MMA warp per query tile
constexpr float kRescaleThreshold = 8.0f;  // lazy O rescale: only when the row max grows by > 2^8
#ifndef FA_FWD_EMU_OF8
#define FA_FWD_EMU_OF8 0
#endif
constexpr int kEmuOf8 = FA_FWD_EMU_OF8;  // element pairs (of every 8) exponentiated on the FMA pipe instead of MUFU

template <int D>
struct FwdCfg {
  static constexpr int kStages = (D == 128) ? 4 : 8;
  static constexpr int kTileBytes = 128 * D * 2;  // one Q / K / V tile
  static constexpr int kSubTileBytes = 128 * 128;  // one 64-column (128-byte) swizzled sub-tile
  // 2 Q tiles + K/V ring + one 64-column O staging sub-tile per query tile + alignment slack + barriers
  static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 2 * kSubTileBytes + 1024 + 256;
};
static_assert(FwdCfg<128>::kSmemBytes <= 232448, "forward smem budget");

// number of 64-column softmax steps a query tile starting at local row `row0` must visit
__device__ __forceinline__ int fwd_num_steps(int row0, const FwdParams& p) {
  if (row0 >= p.n_q) return 0;
  int n = (p.n_kv + kStep - 1) / kStep;
  if (p.causal) {
    const long long last_visible = static_cast<long long>(row0) + kBM - 1 + p.diag;  // for the tile's last row
    if (last_visible < 0) return 0;
    const int nc = static_cast<int>(last_visible / kStep) + 1;
    n = nc < n ? nc : n;
  }
  return n;
}

// One unit of work: a pair of 128-row query tiles of one slice.
struct FwdItem {
  int bh, row0, nt0, nt1, ntmax, n_kv_tiles;
  bool valid;  // false: padding (slice index past the end, or past the last item)
};

// The k-th item of this (persistent) CTA: LPT order over L2-sized slice groups (see the note in ptx.cuh), heaviest
// (latest, under a causal mask) tile pair first, dealt to the CTAs in snake order -- round k runs left to right for
// even k and right to left for odd k -- so every CTA's heavier items are paired with lighter ones.
__device__ __forceinline__ FwdItem fwd_item(const FwdParams& p, int k) {
  const long long G = gridDim.x;
  const long long w = k * G + ((k & 1) ? (G - 1 - static_cast<long long>(blockIdx.x)) : static_cast<long long>(blockIdx.x));
  FwdItem it;
  it.valid = w < p.n_items;
  const unsigned t = static_cast<unsigned>(w >> p.group_log2);
  const unsigned z = t / static_cast<unsigned>(p.npairs);
  const int rank = static_cast<int>(t - z * static_cast<unsigned>(p.npairs));
  it.bh = static_cast<int>((z << p.group_log2) + (static_cast<unsigned>(w) & ((1u << p.group_log2) - 1u)));
  if (it.bh >= p.bh) it.valid = false;
  it.row0 = (p.npairs - 1 - rank) * 2 * kBM;
  it.nt0 = it.valid ? fwd_num_steps(it.row0, p) : 0;
  it.nt1 = it.valid ? fwd_num_steps(it.row0 + kBM, p) : 0;
  it.ntmax = it.nt0 > it.nt1 ? it.nt0 : it.nt1;  // steps
  it.n_kv_tiles = (it.ntmax + 1) >> 1;           // 128-row K/V tiles to stream
  return it;
}

// Persistent kernel: gridDim.x CTAs (one per SM, minus an optional margin left to communication kernels) walk their
// items back to back.  TMEM, barriers and the K/V ring are set up once and simply keep running across item boundaries:
// the ring slot count `gslot` and each query tile's step count `gs` give stage / buffer and phase, so the next item's
// K/V tiles are already streaming in, and its Q tile is loaded and its first two S products issued, while the softmax
// warpgroup still normalises and writes out the previous item's O (through its own 16 KiB staging sub-tile, so the Q
// buffers belong to the loads).
template <int D, bool kBF16>
__global__ void __launch_bounds__(kFwdThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const FwdParams p) {
  using Cfg = FwdCfg<D>;
  constexpr int NS = Cfg::kStages;
  constexpr int kSub = Cfg::kSubTileBytes;
  constexpr int kChunks = D / 64;  // 64-column TMA boxes per tile row
  static_assert((NS & (NS - 1)) == 0, "ring depth must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                            // 2 tiles
  uint8_t* kv_smem = smem + 2 * Cfg::kTileBytes;     // NS tiles
  uint8_t* o_stage = kv_smem + NS * Cfg::kTileBytes; // 2 x 16 KiB: one 64-column O sub-tile per query tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_stage + 2 * kSub);
  uint64_t* q_full = bars;             // [2]     producer -> MMA: Q_i of the item has landed
  uint64_t* s_full = bars + 2;         // [2 tiles][2 buffers]  MMA -> softmax: S_i(step) is in TMEM
  uint64_t* p_ready = bars + 6;        // [2][2]  softmax -> MMA: P_i(step) stored (and O_i rescaled)
  uint64_t* pv_done = bars + 10;       // [2][2]  MMA -> softmax: O_i += P_i(step) V finished (indexed by step parity)
  uint64_t* q_empty = bars + 14;       // [2]     MMA -> producer: the item's last S_i product has read Q_i
  uint64_t* o_free = bars + 16;        // [2]     softmax -> MMA: O_i of the item has been read out of TMEM
  uint64_t* kv_full = bars + 18;       // [NS]
  uint64_t* kv_empty = bars + 18 + NS; // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18 + 2 * NS);
  // pv_done is split by step parity so that every barrier a warpgroup waits on is at most ONE phase behind what it
  // already knows to be complete: s_full(t) only proves PV(t-2) finished, and a parity wait cannot tell "two phases
  // behind" from "done".

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int n_rounds = static_cast<int>((p.n_items + gridDim.x - 1) / gridDim.x);  // items per CTA (tail: padding)

  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_free[i], 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);  // both MMA warps release every stage
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
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
    // ===================================== TMA producer =====================================
    // Two independent load sequences over the CTA's items: the Q tiles (a tile's buffer frees when the previous item's
    // last S product of that tile has run) and the K/V ring (K_0 V_0 K_1 V_1 ... of item after item, a stage frees
    // when both MMA warps have released it).  Polling both keeps the ring streaming across item boundaries.
    if (lane == 0) {
      int q_round = 0, q_tile = 0, q_cnt = 0;  // q_cnt: valid items whose Q has been issued completely
      FwdItem q_item = fwd_item(p, 0);
      int kv_round = 0, kv_t = 0;
      uint32_t gslot = 0;
      FwdItem kv_item = q_item;
      auto skip_q = [&]() {
        while (q_round < n_rounds && !q_item.valid) {
          ++q_round;
          if (q_round < n_rounds) q_item = fwd_item(p, q_round);
        }
      };
      auto skip_kv = [&]() {
        while (kv_round < n_rounds && kv_item.n_kv_tiles == 0) {
          ++kv_round;
          if (kv_round < n_rounds) kv_item = fwd_item(p, kv_round);
        }
      };
      skip_q();
      skip_kv();
      const long long t0 = clock64();
      while (q_round < n_rounds || kv_round < n_rounds) {
        if (q_round < n_rounds && (q_cnt == 0 || mbar_test_wait(&q_empty[q_tile], (q_cnt - 1) & 1))) {
          mbar_arrive_expect_tx(&q_full[q_tile], Cfg::kTileBytes);
          for (int c = 0; c < kChunks; ++c)
            tma_load_3d(q_smem + q_tile * Cfg::kTileBytes + c * kSub, &tm_q, &q_full[q_tile], c * 64,
                        q_item.row0 + q_tile * kBM, q_item.bh);
          if (++q_tile == 2) {
            q_tile = 0;
            ++q_cnt;
            ++q_round;
            if (q_round < n_rounds) q_item = fwd_item(p, q_round);
            skip_q();
          }
        }
        if (kv_round < n_rounds) {
          const uint32_t stage = gslot & (NS - 1);
          if (gslot < NS || mbar_test_wait(&kv_empty[stage], ((gslot / NS) & 1) ^ 1)) {
            mbar_arrive_expect_tx(&kv_full[stage], Cfg::kTileBytes);
            const CUtensorMap* tm = (kv_t & 1) ? &tm_v : &tm_k;
            for (int c = 0; c < kChunks; ++c)
              tma_load_3d(kv_smem + stage * Cfg::kTileBytes + c * kSub, tm, &kv_full[stage], c * 64, (kv_t >> 1) * kBN,
                          kv_item.bh);
            ++gslot;
            if (++kv_t == 2 * kv_item.n_kv_tiles) {
              kv_t = 0;
              ++kv_round;
              if (kv_round < n_rounds) kv_item = fwd_item(p, kv_round);
              skip_kv();
            }
          }
        }
        if (clock64() - t0 > 8 * FA_WAIT_TIMEOUT_CYCLES) {
          printf("fa_sm100 fwd: producer timeout (block %d: q round %d tile %d, kv round %d slot %d of %d rounds)\n",
                 blockIdx.x, q_round, q_tile, kv_round, kv_t, n_rounds);
          __trap();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 9) {
    // ===================================== MMA issuers (warp 9 -> tile 0, warp 10 -> tile 1) =====================
    // Whole warp runs the loop (uniform control flow, descriptors in uniform registers); one elected lane issues.
    // Every warp walks EVERY ring slot (wait full -> use -> release), even past its own tile's last step, so the
    // two-arrival kv_empty barriers stay in lock-step with the producer.
    const int i = warp - 9;
    constexpr uint32_t idesc_s = umma_idesc(kBF16, kBM, kStep, false, false);  // S = Q K^T : A, B K-major
    constexpr uint32_t idesc_o = umma_idesc(kBF16, kBM, D, false, true);       // O += P V : A in TMEM, B MN-major
    constexpr uint32_t kStageLo = Cfg::kTileBytes >> 4;                         // descriptor units per ring stage
    constexpr uint32_t kHalfLo = (kStep * 128) >> 4;                            // second 64 key rows of a tile
    const uint32_t q_lo = umma_desc_lo(smem_u32(q_smem) + i * Cfg::kTileBytes, 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(kv_smem), 16);                 // K as K-major B operand
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(kv_smem), kSub);               // V as MN-major B operand
    const uint32_t t_s = tmem_base + i * kBN;
    const uint32_t t_o = tmem_base + 256 + i * D;

    // `half` = which 64 key rows of the K/V tile, `buf` = which of the tile's two S buffers (running step parity)
    auto issue_s = [&](uint32_t half, uint32_t buf, uint32_t stage) {
      const uint32_t b_lo = k_lo0 + stage * kStageLo + half * kHalfLo;
      const uint32_t d_tmem = t_s + buf * kStep;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        constexpr uint32_t kSubLo = Cfg::kSubTileBytes >> 4;
        const uint32_t off = (kk >> 2) * kSubLo + (kk & 3) * 2;  // 16 elements = 32 B inside the 128-B swizzle row
        umma_ss(d_tmem, umma_desc(q_lo + off), umma_desc(b_lo + off), idesc_s, kk > 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](uint32_t half, uint32_t buf, uint32_t stage, bool acc) {  // O_i += P_i(step) V[half]
      const uint32_t b_lo = v_lo0 + stage * kStageLo + half * kHalfLo;
      const uint32_t a_tmem = t_s + buf * kStep;
#pragma unroll
      for (int kk = 0; kk < kStep / 16; ++kk)  // A: 16 key columns = 8 TMEM columns; B: 16 key rows = 2 KiB
        umma_ts(t_o, a_tmem + kk * 8, umma_desc(b_lo + kk * 128), idesc_o, (acc || kk > 0) ? 1u : 0u);
    };
    auto stage_of = [&](uint32_t slot) { return slot & (NS - 1); };
    auto phase_of = [&](uint32_t slot) { return (slot / NS) & 1u; };

    uint32_t slot0 = 0;  // ring slot of this item's K tile 0 (K tile j -> slot0 + 2j, V tile j -> slot0 + 2j + 1)
    int gs0 = 0;         // this tile's softmax steps before the current item
    int cnt = 0;         // valid items before the current one
    for (int round = 0; round < n_rounds; ++round) {
      const FwdItem item = fwd_item(p, round);
      if (!item.valid) continue;
      const int nti = i == 0 ? item.nt0 : item.nt1;
      const int ntmax = item.ntmax;
      mbar_wait(&q_full[i], cnt & 1);
      if (ntmax > 0) {
        mbar_wait(&kv_full[stage_of(slot0)], phase_of(slot0));
        tc_fence_after();
        if (elect_one()) {
          if (0 < nti) {
            issue_s(0, gs0 & 1, stage_of(slot0));
            tc_commit(&s_full[i * 2 + (gs0 & 1)]);
          }
          if (1 < nti) {
            issue_s(1, (gs0 + 1) & 1, stage_of(slot0));
            tc_commit(&s_full[i * 2 + ((gs0 + 1) & 1)]);
          }
          if (nti <= 2) tc_commit(&q_empty[i]);  // no further S product reads Q_i in this item
          tc_commit(&kv_empty[stage_of(slot0)]);  // K tile 0 only feeds steps 0 and 1
        }
        __syncwarp();

        for (int t = 0; t < ntmax; ++t) {
          const uint32_t sv = slot0 + 2 * (t >> 1) + 1;  // ring slot of the V tile of step t
          const int s2 = t + 2;                          // the S step issued in this iteration
          const uint32_t sk = slot0 + 2 * (s2 >> 1);     // ring slot of its K tile
          if ((t & 1) == 0) mbar_wait(&kv_full[stage_of(sv)], phase_of(sv));
          if ((s2 & 1) == 0 && s2 < ntmax) mbar_wait(&kv_full[stage_of(sk)], phase_of(sk));
          if (t < nti) {
            const int g = gs0 + t;
            mbar_wait(&p_ready[i * 2 + (g & 1)], (g >> 1) & 1);
            if (t == 0 && cnt > 0) mbar_wait(&o_free[i], (cnt - 1) & 1);  // the previous item's O_i has been read out
          }
          tc_fence_after();
          if (elect_one()) {
            if (t < nti) {
              const int g = gs0 + t;
              issue_pv(t & 1, g & 1, stage_of(sv), t > 0);
              tc_commit(&pv_done[i * 2 + (g & 1)]);
            }
            if (s2 < nti) {
              const int g = gs0 + s2;
              issue_s(s2 & 1, g & 1, stage_of(sk));
              tc_commit(&s_full[i * 2 + (g & 1)]);
              if (s2 == nti - 1) tc_commit(&q_empty[i]);
            }
            // V tile: released after its second half (or the very last step); K tile: after its odd (or last) S step
            if ((t & 1) == 1 || t == ntmax - 1) tc_commit(&kv_empty[stage_of(sv)]);
            if (s2 <= ntmax - 1 && ((s2 & 1) == 1 || s2 == ntmax - 1)) tc_commit(&kv_empty[stage_of(sk)]);
          }
          __syncwarp();
        }
      } else {
        tc_commit_elect(&q_empty[i]);  // nothing visible: hand the Q buffer straight back
      }
      if (nti == 0 && cnt > 0) mbar_wait(&o_free[i], (cnt - 1) & 1);  // keep the per-item phases in lock-step
      slot0 += 2 * item.n_kv_tiles;
      gs0 += nti;
      ++cnt;
    }
  } else {
    // ===================================== softmax warpgroups =====================================
    const int wg = warp >> 2;            // query tile 0 / 1
    const int row = threadIdx.x & 127;   // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_sel + wg * kBN;
    const uint32_t t_o = tmem_base + lane_sel + 256 + wg * D;
    const float c = p.scale_log2;
    uint8_t* stage_tile = o_stage + wg * kSub;

    int gs0 = 0;
    for (int round = 0; round < n_rounds; ++round) {
      const FwdItem item = fwd_item(p, round);
      if (!item.valid) continue;
      const int bh = item.bh;
      const int nt = wg == 0 ? item.nt0 : item.nt1;
      const int tile_row0 = item.row0 + wg * kBM;
      const int row_l = tile_row0 + row;   // row inside this slice
      // largest visible key index for this row, in slice-local coordinates (fits an int: n_kv, |diag| <= 2^30, and a
      // row that sees nothing is clamped to -1)
      int vis = p.n_kv - 1;
      if (p.causal) {
        const long long cv = static_cast<long long>(row_l) + p.diag;
        vis = cv < vis ? static_cast<int>(cv < -1 ? -1 : cv) : vis;
      }

      float m_ref = -INFINITY;  // reference max (raw score units) all stored exponentials are relative to
      float l_sum = 0.f;

      for (int j = 0; j < nt; ++j) {  // j = 64-column step
        const int g = gs0 + j;
        const int buf = g & 1;
        const uint32_t t_sb = t_s + buf * kStep;
        mbar_wait(&s_full[wg * 2 + buf], (g >> 1) & 1);
        tc_fence_after();
        float s[kStep];
        tmem_ld32(t_sb, reinterpret_cast<uint32_t*>(s));
        tmem_ld32(t_sb + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tc_wait_ld();

        const int lim = vis - j * kStep;  // last visible column of this step (may be negative: nothing visible)
        if (lim < kStep - 1) {
#pragma unroll
          for (int x = 0; x < kStep; ++x) s[x] = (x > lim) ? -INFINITY : s[x];
        }

        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int x = 4; x < kStep; x += 4) {
          mx0 = fmaxf(mx0, s[x]);
          mx1 = fmaxf(mx1, s[x + 1]);
          mx2 = fmaxf(mx2, s[x + 2]);
          mx3 = fmaxf(mx3, s[x + 3]);
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float m_new = fmaxf(m_ref, m_tile);

        bool rescale = false;
        float alpha = 1.f;
        if (j == 0) {
          m_ref = m_new;
        } else {
          const bool need = (m_new - m_ref) * c > kRescaleThreshold;  // (-inf -> finite) is "needed"; NaN is not
          if (__any_sync(0xffffffffu, need)) {
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            alpha = ex2((m_ref - m_safe) * c);
            l_sum *= alpha;
            m_ref = m_new;
            rescale = true;
          }
        }
        const float mc = ((m_ref == -INFINITY) ? 0.f : m_ref) * c;

        // P = 2^(S*c - m*c), row-sum in fp32, stored 16-bit into the first 32 columns of this S buffer.
        // MUFU (16 ex2/clk/SM) is the co-bottleneck of the tensor pipe at d=128, so kEmuOf8 of every 8 element pairs
        // take the polynomial path on the FMA pipe instead; all arithmetic is packed fp32x2.
        float2 ls_a = make_float2(0.f, 0.f), ls_b = make_float2(0.f, 0.f);
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
#pragma unroll
        for (int q2 = 0; q2 < kStep / 32; ++q2) {
          uint32_t pk[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float2 t = ffma2(make_float2(s[q2 * 32 + 2 * x], s[q2 * 32 + 2 * x + 1]), c2, nmc2);
            float2 pv;
            if ((x & 7) < kEmuOf8) {
              pv = ex2_poly2(t);
            } else {
              pv.x = ex2(t.x);
              pv.y = ex2(t.y);
            }
            if (x & 1) ls_b = fadd2(ls_b, pv); else ls_a = fadd2(ls_a, pv);
            pk[x] = pack2<kBF16>(pv.x, pv.y);
          }
          tmem_st16(t_sb + q2 * 16, pk);
        }
        l_sum += (ls_a.x + ls_a.y) + (ls_b.x + ls_b.y);

        if (rescale) {  // warp-uniform
          mbar_wait(&pv_done[wg * 2 + ((g - 1) & 1)], ((g - 1) >> 1) & 1);
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

      // ------------------------------- item epilogue: O / l, lse, optional LSE merge, TMA store -------------------------------
      // The MMA warp and the producer are already on the next item (Q load, first S products); only the first P V of
      // the next item waits for `o_free`, signalled below right after the last read of O from TMEM.
      if (nt > 0) {
        const int g1 = gs0 + nt - 1;
        if (nt > 1) mbar_wait(&pv_done[wg * 2 + ((g1 - 1) & 1)], ((g1 - 1) >> 1) & 1);
        mbar_wait(&pv_done[wg * 2 + (g1 & 1)], (g1 >> 1) & 1);
        tc_fence_after();
      }
      const bool has_mass = l_sum > 0.f;
      float w_cur = has_mass ? 1.f / l_sum : 0.f;
      const float m_fin = (m_ref == -INFINITY) ? 0.f : m_ref;
      float lse_val = has_mass ? (m_fin * c + log2f(l_sum)) * 0.6931471805599453f : -INFINITY;
      float w_prev = 0.f;
      const bool merge = (p.lse_prev != nullptr) && (row_l < p.n_q);
      if (merge) {
        const float lp = p.lse_prev[static_cast<long long>(bh) * p.lse_bh_stride + row_l];
        const float hi = fmaxf(lp, lse_val);
        if (hi == -INFINITY) {
          w_prev = 0.f;
          w_cur = 0.f;
        } else {
          const float e_prev = __expf(lp - hi), e_cur = __expf(lse_val - hi);
          const float tot = e_prev + e_cur;
          w_prev = e_prev / tot;
          w_cur *= e_cur / tot;
          lse_val = hi + __logf(tot);
        }
      }
      const uint32_t* o_prev_row =
          merge ? reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(p.o_prev) +
                                                    static_cast<long long>(bh) * p.o_bh_stride +
                                                    static_cast<long long>(row_l) * p.d)
                : nullptr;
      if (row_l < p.n_q) p.lse[static_cast<long long>(bh) * p.lse_bh_stride + row_l] = lse_val;
#pragma unroll
      for (int half = 0; half < kChunks; ++half) {  // 64 output columns per round through the 16 KiB staging sub-tile
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          float o[32];
          if (nt > 0) {
            tmem_ld32(t_o + half * 64 + q4 * 32, reinterpret_cast<uint32_t*>(o));
            tc_wait_ld();
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x) o[x] = 0.f;
          }
          if (half == kChunks - 1 && q4 == 1) {  // last read of O: the next item's first P V may overwrite it
            tc_fence_before();
            mbar_arrive(&o_free[wg]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            float a = o[2 * x] * w_cur, b = o[2 * x + 1] * w_cur;
            if (merge && half * 64 + q4 * 32 + 2 * x < p.d) {
              const float2 pv = unpack2<kBF16>(o_prev_row[half * 32 + q4 * 16 + x]);
              a = fmaf(pv.x, w_prev, a);
              b = fmaf(pv.y, w_prev, b);
            }
            pk[x] = pack2<kBF16>(a, b);
          }
          // 32 columns = 64 B = four 16-byte chunks of the 128-byte swizzled row
          uint8_t* sub = stage_tile + row * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int chunk = q4 * 4 + ch;
            *reinterpret_cast<uint4*>(sub + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + wg, 128);
        if (row == 0) {
          if (tile_row0 < p.n_q) tma_store_3d(&tm_o, stage_tile, half * 64, tile_row0, bh);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        named_bar_sync(1 + wg, 128);  // the staging sub-tile is free again (next round / next item)
      }
      gs0 += nt;
    }
    if (row == 0) tma_store_wait_exit();  // the stores themselves complete by grid end
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

template <int D, bool kBF16>
static int launch_fwd(const Geometry& g, const void* q, const void* k, const void* v, void* o, float* lse,
                      const void* o_prev, const float* lse_prev, cudaStream_t stream) {
  using Cfg = FwdCfg<D>;
  const int elem = kBF16 ? kElemBF16 : kElemF16;
  CUtensorMap tm_q, tm_k, tm_v, tm_o;
  int rc;
  if ((rc = make_tmap_3d(&tm_q, q, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kBM))) return rc;
  if ((rc = make_tmap_3d(&tm_k, k, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kBN))) return rc;
  if ((rc = make_tmap_3d(&tm_v, v, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kBN))) return rc;
  if ((rc = make_tmap_3d(&tm_o, o, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kBM))) return rc;

  FwdParams p;
  p.lse = lse;
  p.o_prev = o_prev;
  p.lse_prev = lse_prev;
  p.lse_bh_stride = g.lse_bh_stride;
  p.o_bh_stride = g.q_bh_stride;
  p.n_q = static_cast<int>(g.n_q);
  p.n_kv = static_cast<int>(g.n_kv);
  p.bh = static_cast<int>(g.bh);
  p.causal = g.causal;
  p.diag = g.diag;
  p.d = g.d;
  p.npairs = static_cast<int>((g.n_q + 2 * kBM - 1) / (2 * kBM));
  p.group_log2 = sched_group_log2(g.causal != 0, p.npairs, g.bh);
  const long long n_groups = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  p.n_items = (n_groups * p.npairs) << p.group_log2;
  p.scale_log2 = g.scale * 1.4426950408889634f;

  auto kern = fa_fwd_kernel<D, kBF16>;
  static bool attr_set[64];  // per device (function attributes are per context); benign race: idempotent
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long ctas = persistent_ctas(p.n_items);
  if (ctas <= 0) return FA_SM100_EDEVICE;
  const dim3 grid(static_cast<unsigned>(ctas), 1, 1);
  kern<<<grid, kFwdThreads, Cfg::kSmemBytes, stream>>>(tm_q, tm_k, tm_v, tm_o, p);
  return launch_status();
}

}  // namespace fa

extern "C" int fa_sm100_fwd(const fa_sm100_shape* s, const void* q, const void* k, const void* v, void* o,
                            float* lse, const void* o_prev, const float* lse_prev, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g);
  if (rc) return rc;
  if (!fa::aligned16(q) || !fa::aligned16(k) || !fa::aligned16(v) || !fa::aligned16(o) || lse == nullptr)
    return FA_SM100_EINVAL_PTR;
  if ((o_prev == nullptr) != (lse_prev == nullptr)) return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g.dp == 128) {
    return g.dtype == FA_SM100_DTYPE_BF16 ? fa::launch_fwd<128, true>(g, q, k, v, o, lse, o_prev, lse_prev, st)
                                          : fa::launch_fwd<128, false>(g, q, k, v, o, lse, o_prev, lse_prev, st);
  }
  return g.dtype == FA_SM100_DTYPE_BF16 ? fa::launch_fwd<64, true>(g, q, k, v, o, lse, o_prev, lse_prev, st)
                                        : fa::launch_fwd<64, false>(g, q, k, v, o, lse, o_prev, lse_prev, st);
}
