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
  int n_q, n_kv, bh, causal, diag, npairs, group_log2;
  int d;             // true head dim (<= D): row stride of o_prev; columns [d, D) are zero-filled / clipped by TMA
  float scale_log2;  // softmax_scale * log2(e)
  // ---- extended variant only (kExt): block-sparse tile mask and dropout ----
  const uint8_t* block_mask;  // (ceil(n_q/128), ceil(n_kv/128)) per slice (or shared), nonzero = tile is computed; nullable
  long long mask_bh_stride;   // elements between slices' masks (0: one mask for every slice)
  int mask_cols;              // = ceil(n_kv / 128)
  uint32_t seed_lo, seed_hi, rng_offset;
  uint32_t drop_threshold;    // an element is dropped iff its random byte < threshold (0: no dropout)
  float drop_scale;           // 256 / (256 - threshold)
  long long q_row0, kv_col0;  // global offsets: the random bits depend on GLOBAL coordinates (sharding-invariant)
};
constexpr int kMaxListedTiles = 4096;  // extended variant: K/V tiles per slice the compacted tile list can hold

constexpr int kBM = 128;  // query rows per tile
constexpr int kBN = 128;  // key rows per K/V tile (TMA granularity)
constexpr int kStep = 64;  // key columns per softmax step
constexpr int kFwdThreads = 352;  // 8 softmax warps + producer + one MMA warp per query tile
constexpr float kRescaleThreshold = 8.0f;  // lazy O rescale: only when the row max grows by > 2^8
#ifndef FA_FWD_EMU_OF8
#define FA_FWD_EMU_OF8 0
#endif
// Element pairs (of every 8) exponentiated on the FMA pipe (Cody-Waite + degree-3 polynomial, ex2_poly2) instead of
// MUFU.  Off: measured on B200 (profiles/r02s_poly_exp2_fraction.log) 1 of 8 gains 5 % at d = 64 non-causal (849 -> 896
// TFLOP/s at N = 8K, MUFU being the clear limiter there), nothing at d = 64 causal or at d = 128, 2 of 8 and more lose
// everywhere -- and the polynomial's 7.5e-5 relative error shows up in the LSE (7e-5 instead of 1e-6).
template <int D>
constexpr int emu_of8() { return FA_FWD_EMU_OF8; }

template <int D>
struct FwdCfg {
  static constexpr int kStages = (D == 128) ? 4 : 8;
  static constexpr int kTileBytes = 128 * D * 2;  // one Q / K / V tile
  static constexpr int kSubTileBytes = 128 * 128;  // one 64-column (128-byte) swizzled sub-tile
  static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kSmemBytesExt = kSmemBytes + kMaxListedTiles * 2;  // + the compacted K/V tile list
};

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

#ifndef FA_GRID_Y_BITS
#define FA_GRID_Y_BITS 15  // grid.y carries up to 2^15 tile ranks; larger indices fold into grid.x.  Tests build with 2 to
#endif                     // exercise the folding at small sizes (tools/README.md)
constexpr int kRankBitsY = FA_GRID_Y_BITS;
// Optional per-CTA lifetime trace (build with -DFA_FWD_TRACE; tools/fwd_trace.py): SM id, wall-clock entry/exit and the
// cycle stamps of the prologue / main loop / epilogue boundaries of every CTA, to size the fixed cost per work item.
#ifdef FA_FWD_TRACE
#define FA_FWD_TRACE_MAX_CTAS 8192
__device__ long long fa_fwd_trace_buf[FA_FWD_TRACE_MAX_CTAS * 10];
__device__ __forceinline__ long long fa_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FA_FWD_STAMP(slot, value)                                                                     \
  do {                                                                                                \
    const unsigned lin_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);             \
    if (lin_ < FA_FWD_TRACE_MAX_CTAS) fa_fwd_trace_buf[lin_ * 10 + (slot)] = (value);                 \
  } while (0)
#else
#define FA_FWD_STAMP(slot, value) do { } while (0)
#endif

// kExt = false is the dense kernel.  kExt = true adds (semantics of the dense branch of the
// reference's stand-alone module, src/fa3/torch/flashattention_pytorch.py:80-87: masked_fill -> softmax -> dropout -> @V,
// with the tile skip of its block-sparse branch, :124):
//   * a block-sparse mask over 128 x 128 tiles: a K/V tile is streamed only if it is active for at least one of the
//     CTA's two query tiles (a compacted list of such tiles is built in shared memory before the main loop); where it
//     is active for only one of them the other one masks its scores to -inf;
//   * dropout on the normalised probabilities: the row sum uses the un-dropped P (the softmax denominator and the LSE
//     do not depend on dropout), the P handed to the P V product is P * keep / (1 - p), keep from Philox bits.
template <int D, bool kBF16, bool kExt>
__global__ void __launch_bounds__(kFwdThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const FwdParams p) {
  using Cfg = FwdCfg<D>;
  constexpr int NS = Cfg::kStages;
  constexpr int kSub = Cfg::kSubTileBytes;
  constexpr int kChunks = D / 64;  // 64-column TMA boxes per tile row

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                            // 2 tiles
  uint8_t* kv_smem = smem + 2 * Cfg::kTileBytes;     // NS tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + NS * Cfg::kTileBytes);
  uint64_t* q_full = bars;             // [2]
  uint64_t* s_full = bars + 2;         // [2 tiles][2 buffers]  MMA -> softmax: S_i(t) is in TMEM
  uint64_t* p_ready = bars + 6;        // [2][2]  softmax -> MMA: P_i(t) stored (and O_i rescaled)
  uint64_t* pv_done = bars + 10;       // [2][2]  MMA -> softmax: O_i += P_i(t) V finished (indexed by t&1, see below)
  uint64_t* kv_full = bars + 14;       // [NS]
  uint64_t* kv_empty = bars + 14 + NS; // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14 + 2 * NS);
  uint16_t* tile_list = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(bars) + 256);  // kExt only
  int* list_len = reinterpret_cast<int*>(tmem_slot + 1);                                       // kExt only
  // pv_done is split by step parity so that every barrier a warpgroup waits on is at most ONE phase behind what it
  // already knows to be complete: s_full(t) only proves PV(t-2) finished, and a parity wait cannot tell "two phases
  // behind" from "done".

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
#ifdef FA_FWD_TRACE
  if (threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    FA_FWD_STAMP(0, static_cast<long long>(smid));
    FA_FWD_STAMP(1, fa_globaltimer());
    FA_FWD_STAMP(2, clock64());
  }
#endif

  // heavy (late, for causal) tile pairs first, over groups of slices small enough for their K/V to stay L2-resident
  // Work-item order (see the note in ptx.cuh for the idea), expressed through the grid shape so that the kernel
  // needs no division: x = slice inside its group (fastest), y = tile-pair rank (heaviest first), z = slice group.
  // Ranks beyond the y limit of a grid are folded into x above the slice bits.
  const uint32_t rank = ((blockIdx.x >> p.group_log2) << kRankBitsY) + blockIdx.y;
  const int bh = static_cast<int>((blockIdx.z << p.group_log2) + (blockIdx.x & ((1u << p.group_log2) - 1u)));
  if (bh >= p.bh || rank >= static_cast<uint32_t>(p.npairs)) return;  // padding of the last group / folded ranks
  const int pair = p.npairs - 1 - static_cast<int>(rank);
  const int row0_t0 = pair * 2 * kBM;
  int nt0 = fwd_num_steps(row0_t0, p);
  int nt1 = fwd_num_steps(row0_t0 + kBM, p);
  int ntmax = nt0 > nt1 ? nt0 : nt1;   // steps
  int n_kv_tiles = (ntmax + 1) >> 1;   // 128-row K/V tiles to stream
  const bool sparse = kExt && p.block_mask != nullptr;
  // k-th streamed K/V tile -> its index along the sequence (dense: the identity; sparse: the compacted list, whose
  // entries also carry which of the two query tiles the K/V tile is active for, bits 14 and 15)
  auto kv_tile = [&](int k) -> int { return sparse ? (tile_list[k] & 0x0FFF) : k; };
  if constexpr (kExt) {
    if (sparse && warp == 8) {
      // warp 8 compacts the active K/V tiles of this tile pair (inside the causal / length limit computed above)
      const uint8_t* m0 = p.block_mask + static_cast<long long>(bh) * p.mask_bh_stride +
                          static_cast<long long>(row0_t0 / kBM) * p.mask_cols;
      const bool has_t1 = row0_t0 + kBM < p.n_q;
      int n = 0;
      for (int base = 0; base < n_kv_tiles; base += 32) {
        const int jt = base + lane;
        uint32_t f = 0;
        if (jt < n_kv_tiles) {
          // a query tile only counts where it can see the K/V tile at all (its own causal / length limit)
          if (m0[jt] != 0 && jt * 2 < nt0) f |= 1u;
          if (has_t1 && m0[p.mask_cols + jt] != 0 && jt * 2 < nt1) f |= 2u;
        }
        const uint32_t any = __ballot_sync(0xffffffffu, f != 0);
        if (f != 0) tile_list[n + __popc(any & ((1u << lane) - 1u))] = static_cast<uint16_t>(jt | (f << 14));
        n += __popc(any);
      }
      __syncwarp();
      if (lane == 0) *list_len = n;
      __syncwarp();
    }
  }

  // The producer lane initialises the barriers and starts the first loads (Q and up to NS K/V tiles) BEFORE the
  // block-wide sync, so the TMA latency overlaps the TMEM allocation and the rest of the prologue.
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&q_full[i], 1);
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
    for (int i = 0; i < 2; ++i) {
      mbar_arrive_expect_tx(&q_full[i], Cfg::kTileBytes);
      for (int c = 0; c < kChunks; ++c)
        tma_load_3d(q_smem + i * Cfg::kTileBytes + c * kSub, &tm_q, &q_full[i], c * 64, row0_t0 + i * kBM, bh);
    }
    const int first_tiles = sparse ? *list_len : n_kv_tiles;
    for (int t = 0; t < 2 * first_tiles && t < NS; ++t) {
      mbar_arrive_expect_tx(&kv_full[t], Cfg::kTileBytes);
      const CUtensorMap* tm = (t & 1) ? &tm_v : &tm_k;
      for (int c = 0; c < kChunks; ++c)
        tma_load_3d(kv_smem + t * Cfg::kTileBytes + c * kSub, tm, &kv_full[t], c * 64, kv_tile(t >> 1) * kBN, bh);
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
  if (threadIdx.x == 0) FA_FWD_STAMP(3, clock64());  // prologue done (barriers, TMEM, first loads issued)
  if constexpr (kExt) {
    if (sparse) {
      // every listed tile is visited in full by both query tiles (the inactive / invisible side masks itself); only the
      // sequence's ragged last tile keeps its one-step form
      n_kv_tiles = *list_len;
      ntmax = 2 * n_kv_tiles;
      if (n_kv_tiles > 0 && kv_tile(n_kv_tiles - 1) * kBN + kStep >= p.n_kv) ntmax -= 1;
      nt0 = row0_t0 < p.n_q ? ntmax : 0;
      nt1 = row0_t0 + kBM < p.n_q ? ntmax : 0;
    }
  }

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {  // slots 0..NS-1 (and Q) were issued in the prologue
      for (int t = NS; t < 2 * n_kv_tiles; ++t) {
        const int stage = t % NS;
        mbar_wait(&kv_empty[stage], ((t / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::kTileBytes);
        const CUtensorMap* tm = (t & 1) ? &tm_v : &tm_k;
        for (int c = 0; c < kChunks; ++c)
          tma_load_3d(kv_smem + stage * Cfg::kTileBytes + c * kSub, tm, &kv_full[stage], c * 64, kv_tile(t >> 1) * kBN,
                      bh);
      }
    }
    __syncwarp();
  } else if (warp >= 9) {
    // ===================================== MMA issuers (warp 9 -> tile 0, warp 10 -> tile 1) =====================
    // Whole warp runs the loop (uniform control flow, descriptors in uniform registers); one elected lane issues.
    // Every warp walks EVERY ring slot (wait full -> use -> release), even past its own tile's last step, so the
    // two-arrival kv_empty barriers stay in lock-step with the producer.
    const int i = warp - 9;
    const int nti = i == 0 ? nt0 : nt1;
    if (ntmax > 0) {
      constexpr uint32_t idesc_s = umma_idesc(kBF16, kBM, kStep, false, false);  // S = Q K^T : A, B K-major
      constexpr uint32_t idesc_o = umma_idesc(kBF16, kBM, D, false, true);       // O += P V : A in TMEM, B MN-major
      constexpr uint32_t kStageLo = Cfg::kTileBytes >> 4;                         // descriptor units per ring stage
      constexpr uint32_t kHalfLo = (kStep * 128) >> 4;                            // second 64 key rows of a tile
      const uint32_t q_lo = umma_desc_lo(smem_u32(q_smem) + i * Cfg::kTileBytes, 16);
      const uint32_t k_lo0 = umma_desc_lo(smem_u32(kv_smem), 16);                 // K as K-major B operand
      const uint32_t v_lo0 = umma_desc_lo(smem_u32(kv_smem), kSub);               // V as MN-major B operand
      const uint32_t t_s = tmem_base + i * kBN;
      const uint32_t t_o = tmem_base + 256 + i * D;

      auto issue_s = [&](int step, uint32_t stage) {  // S_i(step) into buffer step&1
        const uint32_t b_lo = k_lo0 + stage * kStageLo + (step & 1) * kHalfLo;
        const uint32_t d_tmem = t_s + (step & 1) * kStep;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          constexpr uint32_t kSubLo = Cfg::kSubTileBytes >> 4;
          const uint32_t off = (kk >> 2) * kSubLo + (kk & 3) * 2;  // 16 elements = 32 B inside the 128-B swizzle row
          umma_ss(d_tmem, umma_desc(q_lo + off), umma_desc(b_lo + off), idesc_s, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int step, uint32_t stage, bool acc) {  // O_i += P_i(step) V[half]
        const uint32_t b_lo = v_lo0 + stage * kStageLo + (step & 1) * kHalfLo;
        const uint32_t a_tmem = t_s + (step & 1) * kStep;
#pragma unroll
        for (int kk = 0; kk < kStep / 16; ++kk)  // A: 16 key columns = 8 TMEM columns; B: 16 key rows = 2 KiB
          umma_ts(t_o, a_tmem + kk * 8, umma_desc(b_lo + kk * 128), idesc_o, (acc || kk > 0) ? 1u : 0u);
      };
      // ring slot of K tile j is 2j, of V tile j is 2j+1; NS is a power of two
      static_assert((NS & (NS - 1)) == 0, "ring depth must be a power of two");
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
        tc_commit(&kv_empty[0]);  // K tile 0 only feeds steps 0 and 1
      }
      __syncwarp();

      for (int t = 0; t < ntmax; ++t) {
        const uint32_t sv = 2 * (t >> 1) + 1;  // ring slot of the V tile of step t
        const int s2 = t + 2;                  // the S step issued in this iteration
        const uint32_t sk = 2 * (s2 >> 1);     // ring slot of its K tile
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
          // V tile: released after its second half (or the very last step); K tile: after its odd (or last) S step
          if ((t & 1) == 1 || t == ntmax - 1) tc_commit(&kv_empty[stage_of(sv)]);
          if (s2 <= ntmax - 1 && ((s2 & 1) == 1 || s2 == ntmax - 1)) tc_commit(&kv_empty[stage_of(sk)]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax warpgroups =====================================
    const int wg = warp >> 2;            // query tile 0 / 1
    const int row = threadIdx.x & 127;   // row inside the tile == TMEM lane
    const int nt = wg == 0 ? nt0 : nt1;
    const int tile_row0 = row0_t0 + wg * kBM;
    const int row_l = tile_row0 + row;   // row inside this slice
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_sel + wg * kBN;
    const uint32_t t_o = tmem_base + lane_sel + 256 + wg * D;
    const float c = p.scale_log2;
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
      const int buf = j & 1;
      const uint32_t t_sb = t_s + buf * kStep;
      mbar_wait(&s_full[wg * 2 + buf], (j >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0 && j == 0) FA_FWD_STAMP(4, clock64());  // first scores have arrived
      float s[kStep];
      tmem_ld32(t_sb, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(t_sb + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tc_wait_ld();

      // first key column of this step (sparse: the listed tile's position along the sequence)
      const int col0 = kExt ? kv_tile(j >> 1) * kBN + (j & 1) * kStep : j * kStep;
      int lim = vis - col0;  // last visible column of this step (may be negative: nothing visible)
      if constexpr (kExt) {
        if (sparse && !((tile_list[j >> 1] >> (14 + wg)) & 1)) lim = -1;  // tile not active for this query tile
      }
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
      // MUFU (16 ex2/clk/SM) is the co-bottleneck of the tensor pipe, so emu_of8<D>() of every 8 element pairs take
      // the polynomial path on the FMA pipe instead; all arithmetic is packed fp32x2.
      float2 ls_a = make_float2(0.f, 0.f), ls_b = make_float2(0.f, 0.f);
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
      const bool dropout = kExt && p.drop_threshold != 0;
#pragma unroll
      for (int q2 = 0; q2 < kStep / 32; ++q2) {
        uint32_t pk[16];
        uint32_t rnd[8];  // kExt: one random word per 4 key columns (byte b = column b of the group)
        if constexpr (kExt) {
          if (dropout) {
            // One Philox call covers 4 queries x 4 keys and this thread needs only word (query & 3) of it, so the four
            // lanes of a quad (four consecutive query rows: same query >> 2) share the calls: lane L computes the
            // calls of key groups g with (g & 3) == (L & 3) and the words are exchanged with three xor-shuffles per
            // round -- in sub-round k lane L hands word ((L & 3) ^ k) to lane L ^ k, which is exactly the word that
            // lane's query needs.  2 calls + 6 shuffles per 32 columns instead of 8 calls.
            const uint32_t qg = static_cast<uint32_t>(p.q_row0 + row_l);
            const uint32_t kg = static_cast<uint32_t>(p.kv_col0 + col0 + q2 * 32);
            const uint32_t me = lane & 3;
            auto pick = [](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t i) {
              const uint32_t lo = (i & 1) ? a1 : a0, hi = (i & 1) ? a3 : a2;
              return (i & 2) ? hi : lo;
            };
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
              const Philox4 w = philox4x32_7(qg >> 2, (kg >> 2) + 4 * jq + me, static_cast<uint32_t>(bh), p.rng_offset,
                                             p.seed_lo, p.seed_hi);
              uint32_t got[4];  // got[k] = my word of the call of key group 4 jq + (me ^ k)
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) {
                const uint32_t give = pick(w.w[0], w.w[1], w.w[2], w.w[3], me ^ kx);
                got[kx] = kx == 0 ? give : __shfl_xor_sync(0xffffffffu, give, kx);
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) rnd[4 * jq + g] = pick(got[0], got[1], got[2], got[3], me ^ g);
            }
          }
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const float2 t = ffma2(make_float2(s[q2 * 32 + 2 * x], s[q2 * 32 + 2 * x + 1]), c2, nmc2);
          float2 pv;
          if ((x & 7) < emu_of8<D>()) {
            pv = ex2_poly2(t);
          } else {
            pv.x = ex2(t.x);
            pv.y = ex2(t.y);
          }
          if (x & 1) ls_b = fadd2(ls_b, pv); else ls_a = fadd2(ls_a, pv);  // the row sum never sees dropout
          if constexpr (kExt) {
            if (dropout) {
              const uint32_t w = rnd[x >> 1] >> ((x & 1) * 16);
              pv.x = ((w & 0xFFu) >= p.drop_threshold) ? pv.x * p.drop_scale : 0.f;
              pv.y = (((w >> 8) & 0xFFu) >= p.drop_threshold) ? pv.y * p.drop_scale : 0.f;
            }
          }
          pk[x] = pack2<kBF16>(pv.x, pv.y);
        }
        tmem_st16(t_sb + q2 * 16, pk);
      }
      l_sum += (ls_a.x + ls_a.y) + (ls_b.x + ls_b.y);

      if (rescale) {  // warp-uniform
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

    // ------------------------------- epilogue: O / l, lse, optional LSE merge, TMA store -------------------------------
    if (threadIdx.x == 0) FA_FWD_STAMP(5, clock64());  // last softmax step handed over
    if (nt > 0) {
      if (nt > 1) mbar_wait(&pv_done[wg * 2 + ((nt - 2) & 1)], ((nt - 2) >> 1) & 1);
      mbar_wait(&pv_done[wg * 2 + ((nt - 1) & 1)], ((nt - 1) >> 1) & 1);
      tc_fence_after();
    } else {
      mbar_wait(&q_full[wg], 0);  // the Q buffer doubles as the O staging tile: its TMA load must have landed
    }
    if (threadIdx.x == 0) FA_FWD_STAMP(6, clock64());  // last PV finished
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
    uint8_t* stage_tile = q_smem + wg * Cfg::kTileBytes;
    const uint32_t* o_prev_row =
        merge ? reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(p.o_prev) +
                                                  static_cast<long long>(bh) * p.o_bh_stride +
                                                  static_cast<long long>(row_l) * p.d)
              : nullptr;
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
      for (int x = 0; x < 16; ++x) {
        float a = o[2 * x] * w_cur, b = o[2 * x + 1] * w_cur;
        if (merge && q4 * 32 + 2 * x < p.d) {
          const float2 pv = unpack2<kBF16>(o_prev_row[q4 * 16 + x]);
          a = fmaf(pv.x, w_prev, a);
          b = fmaf(pv.y, w_prev, b);
        }
        pk[x] = pack2<kBF16>(a, b);
      }
      // 32 columns = 64 B = four 16-byte chunks of the 128-byte swizzled row
      uint8_t* sub = stage_tile + (q4 >> 1) * kSub + row * 128;
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
      for (int ch = 0; ch < kChunks; ++ch) tma_store_3d(&tm_o, stage_tile + ch * kSub, ch * 64, tile_row0, bh);
      tma_store_commit();
      tma_store_wait_exit();  // the staging tile has been read; the stores themselves complete by grid end
    }
    if (threadIdx.x == 0) FA_FWD_STAMP(7, clock64());  // tile 0 stored
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
#ifdef FA_FWD_TRACE
  if (threadIdx.x == 0) {
    FA_FWD_STAMP(8, clock64());
    FA_FWD_STAMP(9, fa_globaltimer());
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Head dim 129..256 (the reference's benchmark sweeps --head-dim 64 128 256,
// benchmarks/bench_utils.py:256).  The O accumulator alone takes 256 TMEM columns, so the ping-pong pair of query tiles
// of the main kernel does not fit: here a CTA owns ONE 128-row query tile and streams 64-row K/V tiles (one softmax
// step each), S double-buffered in two 64-column TMEM buffers so S(t+1) is computed while the warpgroup exponentiates
// S(t).  TMEM: S0 [0,64)  S1 [64,128)  O [128,384).  SMEM: Q 64 KiB + 4-stage K/V ring of 32 KiB tiles = 192 KiB.
// Warps 0-3 softmax (thread = query row), warp 4 TMA producer, warp 5 MMA issuer.  Forward only: a backward at this
// head dim needs dK and dV accumulators of 256 columns each -- all of TMEM -- and is not provided.
// ------------------------------------------------------------------------------------------------
constexpr int kFwd256Threads = 192;
constexpr int kBN256 = 64;
struct Fwd256Cfg {
  static constexpr int kD = 256;
  static constexpr int kStages = 4;
  static constexpr int kQBytes = 128 * kD * 2;        // 64 KiB: four 64-column sub-tiles of 16 KiB
  static constexpr int kQSub = 128 * 128;
  static constexpr int kKVBytes = kBN256 * kD * 2;    // 32 KiB: four 64-column sub-tiles of 8 KiB
  static constexpr int kKVSub = kBN256 * 128;
  static constexpr int kSmemBytes = kQBytes + kStages * kKVBytes + 1024 + 256;
};

template <bool kBF16>
__global__ void __launch_bounds__(kFwd256Threads, 1)
fa_fwd256_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const FwdParams p) {
  using Cfg = Fwd256Cfg;
  constexpr int D = Cfg::kD, NS = Cfg::kStages, kChunks = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + Cfg::kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + NS * Cfg::kKVBytes);
  uint64_t* q_full = bars;            // [1]
  uint64_t* s_full = bars + 1;        // [2]
  uint64_t* p_ready = bars + 3;       // [2]
  uint64_t* pv_done = bars + 5;       // [2]  (split by step parity, see the main kernel)
  uint64_t* kv_full = bars + 7;       // [NS]
  uint64_t* kv_empty = bars + 7 + NS; // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7 + 2 * NS);

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  // same grid shape as the main kernel (x = slice in group, y = tile rank heaviest first, z = group); here a rank is
  // one query tile
  const uint32_t rank = ((blockIdx.x >> p.group_log2) << kRankBitsY) + blockIdx.y;
  const int bh = static_cast<int>((blockIdx.z << p.group_log2) + (blockIdx.x & ((1u << p.group_log2) - 1u)));
  if (bh >= p.bh || rank >= static_cast<uint32_t>(p.npairs)) return;
  const int tile_row0 = (p.npairs - 1 - static_cast<int>(rank)) * kBM;
  const int nt = fwd_num_steps(tile_row0, p);  // 64-column steps == K/V tiles to stream

  if (warp == 4 && lane == 0) {
    mbar_init(&q_full[0], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    mbar_arrive_expect_tx(&q_full[0], Cfg::kQBytes);
    for (int c = 0; c < kChunks; ++c) tma_load_3d(q_smem + c * Cfg::kQSub, &tm_q, &q_full[0], c * 64, tile_row0, bh);
    for (int t = 0; t < 2 * nt && t < NS; ++t) {
      mbar_arrive_expect_tx(&kv_full[t], Cfg::kKVBytes);
      const CUtensorMap* tm = (t & 1) ? &tm_v : &tm_k;
      for (int c = 0; c < kChunks; ++c)
        tma_load_3d(kv_smem + t * Cfg::kKVBytes + c * Cfg::kKVSub, tm, &kv_full[t], c * 64, (t >> 1) * kBN256, bh);
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

  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      for (int t = NS; t < 2 * nt; ++t) {
        const int stage = t % NS;
        mbar_wait(&kv_empty[stage], ((t / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::kKVBytes);
        const CUtensorMap* tm = (t & 1) ? &tm_v : &tm_k;
        for (int c = 0; c < kChunks; ++c)
          tma_load_3d(kv_smem + stage * Cfg::kKVBytes + c * Cfg::kKVSub, tm, &kv_full[stage], c * 64, (t >> 1) * kBN256,
                      bh);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================================== MMA issuer =====================================
    if (nt > 0) {
      constexpr uint32_t idesc_s = umma_idesc(kBF16, kBM, kStep, false, false);  // S = Q K^T, N = 64
      constexpr uint32_t idesc_o = umma_idesc(kBF16, kBM, D, false, true);       // O += P V, N = 256, V MN-major
      constexpr uint32_t kStageLo = Cfg::kKVBytes >> 4;
      const uint32_t q_lo = umma_desc_lo(smem_u32(q_smem), 16);
      const uint32_t k_lo0 = umma_desc_lo(smem_u32(kv_smem), 16);
      const uint32_t v_lo0 = umma_desc_lo(smem_u32(kv_smem), Cfg::kKVSub);
      const uint32_t t_o = tmem_base + 128;
      auto stage_of = [&](uint32_t slot) { return slot & (NS - 1); };
      auto phase_of = [&](uint32_t slot) { return (slot / NS) & 1u; };
      auto issue_s = [&](int step, uint32_t stage) {  // S(step) into buffer step & 1
        const uint32_t b_lo = k_lo0 + stage * kStageLo;
        const uint32_t d_tmem = tmem_base + (step & 1) * kStep;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ss(d_tmem, umma_desc(q_lo + (kk >> 2) * (Cfg::kQSub >> 4) + (kk & 3) * 2),
                  umma_desc(b_lo + (kk >> 2) * (Cfg::kKVSub >> 4) + (kk & 3) * 2), idesc_s, kk > 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int step, uint32_t stage, bool acc) {
        const uint32_t b_lo = v_lo0 + stage * kStageLo;
        const uint32_t a_tmem = tmem_base + (step & 1) * kStep;
#pragma unroll
        for (int kk = 0; kk < kStep / 16; ++kk)
          umma_ts(t_o, a_tmem + kk * 8, umma_desc(b_lo + kk * 128), idesc_o, (acc || kk > 0) ? 1u : 0u);
      };
      mbar_wait(&q_full[0], 0);
      for (int s0 = 0; s0 < 2 && s0 < nt; ++s0) {  // S(0), S(1): K tiles in ring slots 0 and 2
        const uint32_t sk = 2 * s0;
        mbar_wait(&kv_full[stage_of(sk)], phase_of(sk));
        tc_fence_after();
        if (elect_one()) {
          issue_s(s0, stage_of(sk));
          tc_commit(&s_full[s0]);
          tc_commit(&kv_empty[stage_of(sk)]);
        }
        __syncwarp();
      }
      for (int t = 0; t < nt; ++t) {
        const uint32_t sv = 2 * t + 1;
        const int s2 = t + 2;
        const uint32_t sk = 2 * s2;
        mbar_wait(&kv_full[stage_of(sv)], phase_of(sv));
        if (s2 < nt) mbar_wait(&kv_full[stage_of(sk)], phase_of(sk));
        mbar_wait(&p_ready[t & 1], (t >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(t, stage_of(sv), t > 0);
          tc_commit(&pv_done[t & 1]);
          tc_commit(&kv_empty[stage_of(sv)]);
          if (s2 < nt) {
            issue_s(s2, stage_of(sk));
            tc_commit(&s_full[s2 & 1]);
            tc_commit(&kv_empty[stage_of(sk)]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax warpgroup =====================================
    const int row = threadIdx.x & 127;
    const int row_l = tile_row0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_sel;
    const uint32_t t_o = tmem_base + lane_sel + 128;
    const float c = p.scale_log2;
    int vis = p.n_kv - 1;
    if (p.causal) {
      const long long cv = static_cast<long long>(row_l) + p.diag;
      vis = cv < vis ? static_cast<int>(cv < -1 ? -1 : cv) : vis;
    }
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int buf = j & 1;
      const uint32_t t_sb = t_s + buf * kStep;
      mbar_wait(&s_full[buf], (j >> 1) & 1);
      tc_fence_after();
      float s[kStep];
      tmem_ld32(t_sb, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(t_sb + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tc_wait_ld();
      const int lim = vis - j * kStep;
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
      const float m_new = fmaxf(m_ref, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      bool rescale = false;
      float alpha = 1.f;
      if (j == 0) {
        m_ref = m_new;
      } else {
        const bool need = (m_new - m_ref) * c > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = ex2((m_ref - m_safe) * c);
          l_sum *= alpha;
          m_ref = m_new;
          rescale = true;
        }
      }
      const float mc = ((m_ref == -INFINITY) ? 0.f : m_ref) * c;
      float2 ls_a = make_float2(0.f, 0.f), ls_b = make_float2(0.f, 0.f);
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
#pragma unroll
      for (int q2 = 0; q2 < kStep / 32; ++q2) {
        uint32_t pk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const float2 t = ffma2(make_float2(s[q2 * 32 + 2 * x], s[q2 * 32 + 2 * x + 1]), c2, nmc2);
          float2 pv;
          pv.x = ex2(t.x);
          pv.y = ex2(t.y);
          if (x & 1) ls_b = fadd2(ls_b, pv); else ls_a = fadd2(ls_a, pv);
          pk[x] = pack2<kBF16>(pv.x, pv.y);
        }
        tmem_st16(t_sb + q2 * 16, pk);
      }
      l_sum += (ls_a.x + ls_a.y) + (ls_b.x + ls_b.y);
      if (rescale) {  // warp-uniform
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
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
      mbar_arrive(&p_ready[buf]);
    }
    // ------------------------------- epilogue -------------------------------
    if (nt > 0) {
      if (nt > 1) mbar_wait(&pv_done[(nt - 2) & 1], ((nt - 2) >> 1) & 1);
      mbar_wait(&pv_done[(nt - 1) & 1], ((nt - 1) >> 1) & 1);
      tc_fence_after();
    } else {
      mbar_wait(&q_full[0], 0);  // the Q buffer doubles as the O staging tile: its TMA load must have landed
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
      for (int x = 0; x < 16; ++x) {
        float a = o[2 * x] * w_cur, b = o[2 * x + 1] * w_cur;
        if (merge && q4 * 32 + 2 * x < p.d) {
          const float2 pv = unpack2<kBF16>(o_prev_row[q4 * 16 + x]);
          a = fmaf(pv.x, w_prev, a);
          b = fmaf(pv.y, w_prev, b);
        }
        pk[x] = pack2<kBF16>(a, b);
      }
      uint8_t* sub = q_smem + (q4 >> 1) * Cfg::kQSub + row * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int chunk = (q4 & 1) * 4 + ch;
        *reinterpret_cast<uint4*>(sub + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
      }
    }
    if (row_l < p.n_q) p.lse[static_cast<long long>(bh) * p.lse_bh_stride + row_l] = lse_val;
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (row == 0 && tile_row0 < p.n_q) {
      for (int ch = 0; ch < kChunks; ++ch)
        if (ch * 64 < p.d) tma_store_3d(&tm_o, q_smem + ch * Cfg::kQSub, ch * 64, tile_row0, bh);
      tma_store_commit();
      tma_store_wait_exit();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

template <bool kBF16>
static int launch_fwd256(const Geometry& g, const void* q, const void* k, const void* v, void* o, float* lse,
                         const void* o_prev, const float* lse_prev, cudaStream_t stream) {
  using Cfg = Fwd256Cfg;
  const int elem = kBF16 ? kElemBF16 : kElemF16;
  CUtensorMap tm_q, tm_k, tm_v, tm_o;
  int rc;
  if ((rc = make_tmap_3d(&tm_q, q, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kBM))) return rc;
  if ((rc = make_tmap_3d(&tm_k, k, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kBN256))) return rc;
  if ((rc = make_tmap_3d(&tm_v, v, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kBN256))) return rc;
  if ((rc = make_tmap_3d(&tm_o, o, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kBM))) return rc;
  FwdParams p = {};
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
  p.npairs = static_cast<int>((g.n_q + kBM - 1) / kBM);  // one query tile per CTA
  p.group_log2 = sched_group_log2(g.causal != 0, p.npairs, g.bh);
  while (((g.bh + (1ll << p.group_log2) - 1) >> p.group_log2) > 65535) ++p.group_log2;
  p.scale_log2 = g.scale * 1.4426950408889634f;
  auto kern = fa_fwd256_kernel<kBF16>;
  static bool attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long rank_lo = p.npairs < (1 << kRankBitsY) ? p.npairs : (1 << kRankBitsY);
  const long long rank_hi = (p.npairs + (1 << kRankBitsY) - 1) >> kRankBitsY;
  const long long gx = rank_hi << p.group_log2, gz = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  if (gx > 0x7fffffffll || gz > 65535) return FA_SM100_EINVAL_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rank_lo), static_cast<unsigned>(gz));
  kern<<<grid, kFwd256Threads, Cfg::kSmemBytes, stream>>>(tm_q, tm_k, tm_v, tm_o, p);
  return launch_status();
}

template <int D, bool kBF16, bool kExt>
static int launch_fwd(const Geometry& g, const ExtArgs& ext, const void* q, const void* k, const void* v, void* o,
                      float* lse, const void* o_prev, const float* lse_prev, cudaStream_t stream) {
  using Cfg = FwdCfg<D>;
  constexpr int kSmem = kExt ? Cfg::kSmemBytesExt : Cfg::kSmemBytes;
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
  while (((g.bh + (1ll << p.group_log2) - 1) >> p.group_log2) > 65535) ++p.group_log2;  // grid.z limit
  p.scale_log2 = g.scale * 1.4426950408889634f;
  p.block_mask = ext.block_mask;
  p.mask_bh_stride = ext.mask_bh_stride;
  p.mask_cols = ext.mask_cols;
  p.seed_lo = ext.seed_lo;
  p.seed_hi = ext.seed_hi;
  p.rng_offset = ext.rng_offset;
  p.drop_threshold = ext.drop_threshold;
  p.drop_scale = ext.drop_scale;
  p.q_row0 = ext.q_row0;
  p.kv_col0 = ext.kv_col0;

  auto kern = fa_fwd_kernel<D, kBF16, kExt>;
  static bool attr_set[64];  // per device (function attributes are per context); benign race: idempotent
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long rank_lo = p.npairs < (1 << kRankBitsY) ? p.npairs : (1 << kRankBitsY);
  const long long rank_hi = (p.npairs + (1 << kRankBitsY) - 1) >> kRankBitsY;
  const long long gx = rank_hi << p.group_log2, gz = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  if (gx > 0x7fffffffll || gz > 65535) return FA_SM100_EINVAL_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rank_lo), static_cast<unsigned>(gz));
  kern<<<grid, kFwdThreads, kSmem, stream>>>(tm_q, tm_k, tm_v, tm_o, p);
  return launch_status();
}

template <bool kExt>
static int fwd_dispatch(const Geometry& g, const ExtArgs& ext, const void* q, const void* k, const void* v, void* o,
                        float* lse, const void* o_prev, const float* lse_prev, cudaStream_t st) {
  const bool bf = g.dtype == FA_SM100_DTYPE_BF16;
  if (g.dp == 256) {
    if (kExt) return FA_SM100_EINVAL_HEADDIM;  // block-sparse / dropout variants stop at head dim 128
    return bf ? launch_fwd256<true>(g, q, k, v, o, lse, o_prev, lse_prev, st)
              : launch_fwd256<false>(g, q, k, v, o, lse, o_prev, lse_prev, st);
  }
  if (g.dp == 128) {
    return bf ? launch_fwd<128, true, kExt>(g, ext, q, k, v, o, lse, o_prev, lse_prev, st)
              : launch_fwd<128, false, kExt>(g, ext, q, k, v, o, lse, o_prev, lse_prev, st);
  }
  return bf ? launch_fwd<64, true, kExt>(g, ext, q, k, v, o, lse, o_prev, lse_prev, st)
            : launch_fwd<64, false, kExt>(g, ext, q, k, v, o, lse, o_prev, lse_prev, st);
}

}  // namespace fa

extern "C" int fa_sm100_fwd(const fa_sm100_shape* s, const void* q, const void* k, const void* v, void* o,
                            float* lse, const void* o_prev, const float* lse_prev, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g, /*max_d=*/256);
  if (rc) return rc;
  if (!fa::aligned16(q) || !fa::aligned16(k) || !fa::aligned16(v) || !fa::aligned16(o) || lse == nullptr)
    return FA_SM100_EINVAL_PTR;
  if ((o_prev == nullptr) != (lse_prev == nullptr)) return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  return fa::fwd_dispatch<false>(g, fa::ExtArgs(), q, k, v, o, lse, o_prev, lse_prev, static_cast<cudaStream_t>(stream));
}

extern "C" int fa_sm100_fwd_ex(const fa_sm100_shape* s, const fa_sm100_ext* ext, const void* q, const void* k,
                               const void* v, void* o, float* lse, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g);
  if (rc) return rc;
  fa::ExtArgs ea;
  if ((rc = fa::check_ext(ext, s, fa::kMaxListedTiles, &ea))) return rc;
  if (!fa::aligned16(q) || !fa::aligned16(k) || !fa::aligned16(v) || !fa::aligned16(o) || lse == nullptr)
    return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ea.block_mask == nullptr && ea.drop_threshold == 0)  // nothing extra asked for: the dense kernel
    return fa::fwd_dispatch<false>(g, ea, q, k, v, o, lse, nullptr, nullptr, st);
  return fa::fwd_dispatch<true>(g, ea, q, k, v, o, lse, nullptr, nullptr, st);
}

#ifdef FA_FWD_TRACE
// debug build only: copy the per-CTA lifetime stamps (10 values per CTA) of the last forward launch to the host
extern "C" int fa_sm100_debug_fwd_trace(long long* host_dst, int n_ctas) {
  if (!host_dst || n_ctas <= 0 || n_ctas > FA_FWD_TRACE_MAX_CTAS) return -1;
  return cudaMemcpyFromSymbol(host_dst, fa::fa_fwd_trace_buf, sizeof(long long) * 10 * n_ctas) == cudaSuccess ? n_ctas : -1;
}
#endif
