// Fused attention backward for sm_100a, KV-outer: one CTA owns a 128-row K/V tile, streams the visible 128-row
// Q / dO tiles past it, keeps dK and dV accumulators in TMEM, and reduce-adds its dQ partials into an fp32 buffer.
//
// Replaces the reference's host-side ATen loops fa{1,2,3}_backward (csrc/fa1/fa1_bwd.cu:70-110; algorithmic twin
// src/fa1/torch/impl.py:70-115, whose causal block rule `skip iff first key > last query` is the one used here):
//   S = Q K^T * scale,  P = exp(S - lse),  dV += P^T dO,  dP = dO V^T,  dS = P o (dP - delta),
//   dQ += dS K * scale,  dK += dS^T Q * scale.
// Everything is computed TRANSPOSED (kv rows on TMEM lanes) so that P^T and dS^T are already in the layout the
// tensor core wants for an A operand read straight from TMEM:
//   S^T  = K  Q^T      (A = K  smem K-major,  B = Q  smem K-major)            -> TMEM ST
//   dP^T = V  dO^T     (A = V  smem K-major,  B = dO smem K-major)            -> TMEM DPT
//   dV  += P^T  dO     (A = P^T  in TMEM,     B = dO smem MN-major)           -> TMEM DV
//   dK  += dS^T Q      (A = dS^T in TMEM,     B = Q  smem MN-major)           -> TMEM DK
//   dQ   = dS   K      (A = dS^T smem read MN-major, B = K smem MN-major)     -> TMEM DPT (aliases dP^T/dS^T)
// TMEM columns: ST [0,128)  DPT [128,256)  DV [256,256+D)  DK [256+D,256+2D).
// Two independent MMA issue streams, one warp each, interleaved by the tensor pipe:
//   stream X (owns the ST columns):   S^T(0) ; for each tile i:  [P(i) ready] dV(i) . S^T(i+1)
//   stream Y (owns the DPT columns):  for each tile i:  [dQ(i-1) drained] dP^T(i) ; [dS(i) ready] dK(i) . dQ(i)
// so neither chain waits for the other's softmax phase (a single in-order issuer made dP^T(i) queue behind dV(i)).
//
// Warps: 0-3 / 4-7 compute warpgroups (thread = kv row; WG0 takes query columns 0-63, WG1 64-127),
//        8-11 dQ drain warpgroup (TMEM -> swizzled smem -> TMA reduce-add, thread = query row),
//        12 TMA producer, 13 MMA stream X, 14 MMA stream Y.
#include "ptx.cuh"
#include "fa_host.cuh"

// Optional per-phase timeline of one CTA (build with -DFA_BWD_TRACE; tools/bwd_trace.py reads it back).  clock64 is
// the SM's cycle counter, so all warps of the traced CTA share one time base.
#ifdef FA_BWD_TRACE
#ifndef FA_BWD_TRACE_ITERS
#define FA_BWD_TRACE_ITERS 64
#endif
#define FA_BWD_TRACE_EVENTS 16
__device__ long long fa_bwd_trace_buf[FA_BWD_TRACE_EVENTS * FA_BWD_TRACE_ITERS];
__device__ int fa_bwd_trace_block = 0;
__device__ __forceinline__ int fa_bwd_lin_block() {
  return static_cast<int>(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}
#define FA_TRACE(ev, it)                                                                       \
  do {                                                                                         \
    if (fa_bwd_lin_block() == fa_bwd_trace_block && (it) < FA_BWD_TRACE_ITERS)                 \
      fa_bwd_trace_buf[(ev) * FA_BWD_TRACE_ITERS + (it)] = clock64();                          \
  } while (0)
// per-CTA lifetime stamps of EVERY CTA (8 values each): SM id, wall-clock entry/exit, cycle stamps of the boundaries
#define FA_BWD_LIFE_MAX_CTAS 16384
__device__ long long fa_bwd_life_buf[FA_BWD_LIFE_MAX_CTAS * 8];
__device__ __forceinline__ long long fa_bwd_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FA_LIFE(slot, value)                                                                   \
  do {                                                                                         \
    const int lin_ = fa_bwd_lin_block();                                                       \
    if (lin_ < FA_BWD_LIFE_MAX_CTAS) fa_bwd_life_buf[lin_ * 8 + (slot)] = (value);             \
  } while (0)
#else
#define FA_TRACE(ev, it) do { } while (0)
#define FA_LIFE(slot, value) do { } while (0)
#endif

namespace fa {

struct BwdParams {
  const float* rowstats;  // (bh, nqt, 2, 128): -lse*log2e then -delta, per 128-row query tile
  float* dq_accum;        // fp32 dQ accumulator (bh, n_q, D) with slice stride dq_bh_stride (elements)
  long long dq_bh_stride;
  int n_q, n_kv, bh, causal, diag, nqt, nkt, group_log2;
  int d;  // true head dim (<= D); the tensor maps zero-fill / clip the columns in [d, D)
  int accum_kv;  // 0: dk / dv written in the input dtype   1: fp32 partials reduce-added into given accumulators
                 // 2: fp32 partials stored (overwriting) -- ring attention adds them to the travelling accumulators
  float scale_log2, scale;
  // ---- extended variant only (kExt): block-sparse tile mask and dropout (see fa_fwd_sm100.cu) ----
  const uint8_t* block_mask;  // (nqt, nkt) per slice or shared, nonzero = tile is computed; nullable
  long long mask_bh_stride;
  uint32_t seed_lo, seed_hi, rng_offset;
  uint32_t drop_threshold;    // an element is dropped iff its random byte < threshold (0: no dropout)
  float drop_scale;
  long long q_row0, kv_col0;  // global offsets for the random bits
};
constexpr int kMaxMaskTiles = 4096;  // extended variant: query tiles per slice the active-tile bitmap can hold

constexpr int kBwdThreads = 480;  // 15 warps (16 x 128 registers does not launch: the register file has no slack)
constexpr int kT = 128;  // tile edge (query rows and kv rows)
#ifndef FA_GRID_Y_BITS
#define FA_GRID_Y_BITS 15  // grid.y carries up to 2^15 kv tiles; larger indices fold into grid.x.  Tests build with 2 to
#endif                     // exercise the folding at small sizes (tools/README.md)
constexpr int kRankBitsY = FA_GRID_Y_BITS;

template <int D>
struct BwdCfg {
  static constexpr int kTileBytes = kT * D * 2;  // Q / K / V / dO tile
  static constexpr int kSub = kT * 128;          // one 64-column swizzled sub-tile (128 rows x 128 B)
  static constexpr int kDsBytes = kT * kT * 2;   // dS^T tile (kv x q), 16-bit
  static constexpr int kDqStageBytes = kT * 32 * 4;  // 128 rows x 32 fp32 columns
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kTileBytes;
  static constexpr int kOffQ = kOffV + kTileBytes;        // 2 stages
  static constexpr int kOffDO = kOffQ + 2 * kTileBytes;   // 2 stages
  static constexpr int kOffDS = kOffDO + 2 * kTileBytes;  // dS^T tile
  // dQ staging (2 x 16 KiB): at D = 128 there is no room left, so it borrows the (dead) dO stage of the tile being
  // drained; at D = 64 it has its own buffers
  static constexpr bool kStageInDO = (D == 128);
  static constexpr int kOffStage = kOffDS + kDsBytes;
  static constexpr int kOffStats = kOffStage + (kStageInDO ? 0 : 2 * kDqStageBytes);  // 2 stages x 1 KiB
  static constexpr int kOffBars = kOffStats + 2 * 1024;
  static constexpr int kSmemBytes = kOffBars + 256;
  static constexpr int kOffBitmap = kSmemBytes;  // extended variant: active query tiles of this K/V tile, 1 bit each
  static constexpr int kSmemBytesExt = kOffBitmap + kMaxMaskTiles / 8 + 16;
};
static_assert(BwdCfg<128>::kSmemBytesExt <= 232448, "backward smem budget");

enum BwdBar : int {
  kBarKV = 0, kBarQFull0, kBarQFull1, kBarQEmpty0, kBarQEmpty1, kBarDOFull0, kBarDOFull1, kBarDOEmpty0, kBarDOEmpty1,
  kBarSFull, kBarDPFull, kBarDPIssued, kBarPReady, kBarDSHalf, kBarDSReady, kBarDQFull, kBarDQDrained, kBarStageFree0, kBarStageFree1,
  kBarDKVDone, kBarCount
};

// kExt = true: block-sparse tile mask (inactive query tiles of this K/V tile are skipped by every role) and dropout
// (same Philox bits as the forward); kExt = false is the dense kernel.
template <int D, bool kBF16, bool kExt>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
              const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
              const __grid_constant__ CUtensorMap tm_dv, const BwdParams p) {
  using Cfg = BwdCfg<D>;
  constexpr int kSub = Cfg::kSub;
  constexpr int kChunks = D / 64;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* k_smem = smem + Cfg::kOffK;
  uint8_t* v_smem = smem + Cfg::kOffV;
  uint8_t* q_smem = smem + Cfg::kOffQ;
  uint8_t* do_smem = smem + Cfg::kOffDO;
  uint8_t* ds_smem = smem + Cfg::kOffDS;
  float* stats_smem = reinterpret_cast<float*>(smem + Cfg::kOffStats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBarCount);

  const int warp = static_cast<int>(warp_uniform(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;

  // Work-item order through the grid shape (see the note on work-item order in ptx.cuh; no division in the kernel): x = slice inside
  // its group (fastest), y = kv tile j (ascending = heaviest first under the causal mask), z = slice group; tile
  // indices beyond the y limit of a grid are folded into x above the slice bits.
  const int j = static_cast<int>(((blockIdx.x >> p.group_log2) << kRankBitsY) + blockIdx.y);
  const int bh = static_cast<int>((blockIdx.z << p.group_log2) + (blockIdx.x & ((1u << p.group_log2) - 1u)));
  if (bh >= p.bh || j >= p.nkt) return;  // padding of the last group / folded tile indices
#ifdef FA_BWD_TRACE
  if (threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    FA_LIFE(0, static_cast<long long>(smid));
    FA_LIFE(1, fa_bwd_globaltimer());
    FA_LIFE(2, clock64());
  }
#endif
  int i_min = 0;
  if (p.causal) {
    const int first = j * kT - p.diag;  // first query row that sees this tile's first key
    i_min = first > 0 ? first / kT : 0;
  }
  int n_iter = p.nqt > i_min ? p.nqt - i_min : 0;
  // Query tiles are walked in ascending order.  Dense: i_min, i_min + 1, ...  Sparse (kExt with a block mask): the
  // set bits of a bitmap built below; every role keeps its own cursor and steps it with next_tile().
  const bool sparse = kExt && p.block_mask != nullptr;
  uint32_t* act_bits = reinterpret_cast<uint32_t*>(smem + Cfg::kOffBitmap);  // kExt only
  int* act_count = reinterpret_cast<int*>(smem + Cfg::kOffBitmap + kMaxMaskTiles / 8);
  auto next_tile = [&](int i) -> int {  // smallest active tile index > i (p.nqt if none)
    if (!sparse) return i + 1;
    ++i;
    while (i < p.nqt) {
      const uint32_t w = act_bits[i >> 5] >> (i & 31);
      if (w) return i + __ffs(w) - 1;
      i = (i | 31) + 1;
    }
    return p.nqt;
  };
  if constexpr (kExt) {
    if (sparse && warp == 12) {  // the producer warp builds the bitmap before anything is loaded
      const uint8_t* col = p.block_mask + static_cast<long long>(bh) * p.mask_bh_stride + j;
      int cnt = 0;
      for (int base = 0; base < p.nqt; base += 32) {
        const int i = base + lane;
        const bool on = i >= i_min && i < p.nqt && col[static_cast<long long>(i) * p.nkt] != 0;
        const uint32_t bits = __ballot_sync(0xffffffffu, on);
        if (lane == 0) act_bits[base >> 5] = bits;
        cnt += __popc(bits);
      }
      if (lane == 0) *act_count = cnt;
      __syncwarp();
      n_iter = cnt;
    }
  }
  int first_tile = next_tile(i_min - 1);  // sparse: only the producer warp may trust the bitmap before the block sync

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("fa_sm100 bwd: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  // loads of one query tile (Q + row statistics on q_full, dO on do_full)
  auto issue_q_tile = [&](int it, int i) {
    const int st = it & 1;
    mbar_arrive_expect_tx(&bars[kBarQFull0 + st], Cfg::kTileBytes + 1024);
    for (int c = 0; c < kChunks; ++c)
      tma_load_3d(q_smem + st * Cfg::kTileBytes + c * kSub, &tm_q, &bars[kBarQFull0 + st], c * 64, i * kT, bh);
    bulk_load_1d(stats_smem + st * 256, p.rowstats + (static_cast<long long>(bh) * p.nqt + i) * 256, 1024,
                 &bars[kBarQFull0 + st]);
  };
  auto issue_do_tile = [&](int it, int i) {
    const int st = it & 1;
    mbar_arrive_expect_tx(&bars[kBarDOFull0 + st], Cfg::kTileBytes);
    for (int c = 0; c < kChunks; ++c)
      tma_load_3d(do_smem + st * Cfg::kTileBytes + c * kSub, &tm_do, &bars[kBarDOFull0 + st], c * 64, i * kT, bh);
  };
  // The producer lane initialises the barriers and starts K, V and the first two query tiles BEFORE the block-wide
  // sync, so their TMA latency overlaps the TMEM allocation and the rest of the prologue.
  if (warp == 12 && lane == 0) {
    for (int b = 0; b < kBarCount; ++b) {
      uint32_t count = 1u;
      if (b == kBarPReady || b == kBarDSReady || b == kBarDSHalf) count = 256u;
      if (b == kBarDQDrained) count = 128u;
      // operands shared by both MMA streams are released by two commits
      if (b == kBarQEmpty0 || b == kBarQEmpty1 || b == kBarDOEmpty0 || b == kBarDOEmpty1 || b == kBarDKVDone)
        count = 2u;
      mbar_init(&bars[b], count);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_arrive_expect_tx(&bars[kBarKV], 2 * Cfg::kTileBytes);
    for (int c = 0; c < kChunks; ++c) {
      tma_load_3d(k_smem + c * kSub, &tm_k, &bars[kBarKV], c * 64, j * kT, bh);
      tma_load_3d(v_smem + c * kSub, &tm_v, &bars[kBarKV], c * 64, j * kT, bh);
    }
    for (int it = 0, i = first_tile; it < n_iter && it < 2; ++it, i = next_tile(i)) {
      issue_q_tile(it, i);
      issue_do_tile(it, i);
    }
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 256 + D;
  if constexpr (kExt) {
    if (sparse) {  // bitmap and count were written by the producer warp before the block-wide sync
      n_iter = *act_count;
      first_tile = next_tile(i_min - 1);
    }
  }

  if (warp == 12) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {  // K, V and query tiles 0, 1 were issued in the prologue
      // Q and dO stages free up at different moments (Q(it) after dK(it); dO(it) only after the dQ(it) reduce staged in
      // it has been read), so the two load sequences advance independently: polling both keeps a late dO stage from
      // holding back the next Q tile, which heads the following iteration's critical path.
      int q_next = 2, do_next = 2;
      int q_tile = next_tile(next_tile(first_tile)), do_tile = q_tile;  // third active tile (if any)
      const long long t0 = clock64();
      while (q_next < n_iter || do_next < n_iter) {
        if (q_next < n_iter && mbar_test_wait(&bars[kBarQEmpty0 + (q_next & 1)], ((q_next >> 1) & 1) ^ 1)) {
          issue_q_tile(q_next, q_tile);
          FA_TRACE(12, q_next);
          ++q_next;
          q_tile = next_tile(q_tile);
        }
        if (do_next < n_iter) {
          const uint32_t par = ((do_next >> 1) & 1) ^ 1;
          bool ok = mbar_test_wait(&bars[kBarDOEmpty0 + (do_next & 1)], par);  // dV / dP of the previous user are done
          if constexpr (Cfg::kStageInDO)                                        // ... and so is the dQ reduce staged there
            ok = ok && mbar_test_wait(&bars[kBarStageFree0 + (do_next & 1)], par);
          if (ok) {
            issue_do_tile(do_next, do_tile);
            FA_TRACE(13, do_next);
            ++do_next;
            do_tile = next_tile(do_tile);
          }
        }
        if (clock64() - t0 > FA_WAIT_TIMEOUT_CYCLES) {
          printf("fa_sm100 bwd: producer timeout (block %d,%d,%d q_next %d, do_next %d of %d)\n", blockIdx.x, blockIdx.y,
                 blockIdx.z, q_next, do_next,
                 n_iter);
          __trap();
        }
      }
    }
    __syncwarp();
  } else if (warp == 13 || warp == 14) {
    // ===================================== MMA issuers (uniform control flow, one elected lane) ==================
    if (n_iter > 0) {
      constexpr uint32_t idesc_s = umma_idesc(kBF16, kT, kT, false, false);
      constexpr uint32_t idesc_acc = umma_idesc(kBF16, kT, D, false, true);
      constexpr uint32_t idesc_dq = umma_idesc(kBF16, kT, D, true, true);
      constexpr uint32_t kSubLo = kSub >> 4, kTileLo = Cfg::kTileBytes >> 4;
      // low descriptor words: K-major views (LBO unused = 16 B) and MN-major views (LBO = next 64-column sub-tile)
      const uint32_t k_km = umma_desc_lo(smem_u32(k_smem), 16), v_km = umma_desc_lo(smem_u32(v_smem), 16);
      const uint32_t q_km = umma_desc_lo(smem_u32(q_smem), 16), do_km = umma_desc_lo(smem_u32(do_smem), 16);
      const uint32_t k_mn = umma_desc_lo(smem_u32(k_smem), kSub), q_mn = umma_desc_lo(smem_u32(q_smem), kSub);
      const uint32_t do_mn = umma_desc_lo(smem_u32(do_smem), kSub), ds_mn = umma_desc_lo(smem_u32(ds_smem), kT * 128);

      // D[kv, q] = A[kv, :] . B[q, :]   (both K-major, contraction over the head dim)
      auto mma_kmajor = [&](uint32_t d_col, uint32_t a_lo, uint32_t b_lo) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kSubLo + (kk & 3) * 2;
          umma_ss(tmem_base + d_col, umma_desc(a_lo + off), umma_desc(b_lo + off), idesc_s, kk > 0 ? 1u : 0u);
        }
      };
      // D[kv, d] (+)= A^T-in-TMEM[kv, q] . B[q, d]   (contraction over the 128 query rows; B MN-major)
      // the 16-bit A operand sits in columns [0,32) (queries 0-63, written by WG0) and [64,96) (queries 64-127, WG1)
      auto mma_from_tmem = [&](uint32_t d_col, uint32_t a_col, uint32_t b_lo, bool acc) {
#pragma unroll
        for (int kk = 0; kk < kT / 16; ++kk) {
          const uint32_t a = tmem_base + a_col + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8);
          umma_ts(tmem_base + d_col, a, umma_desc(b_lo + kk * 128), idesc_acc, (acc || kk > 0) ? 1u : 0u);
        }
      };
      // same product restricted to the first (part 0) or second (part 1) 32 queries of each warpgroup's 64
      auto mma_from_tmem_part = [&](uint32_t d_col, uint32_t a_col, uint32_t b_lo, bool acc, int part) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int kk = (x >> 1) * 4 + part * 2 + (x & 1);  // part 0: 0,1,4,5   part 1: 2,3,6,7
          const uint32_t a = tmem_base + a_col + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8);
          umma_ts(tmem_base + d_col, a, umma_desc(b_lo + kk * 128), idesc_acc, (acc || x > 0) ? 1u : 0u);
        }
      };
      // dQ[q, d] = dS[q, kv] . K[kv, d]   (contraction over the 128 kv rows; both operands MN-major)
      auto mma_dq = [&]() {
#pragma unroll
        for (int kk = 0; kk < kT / 16; ++kk)
          umma_ss(tmem_base + kColDPT, umma_desc(ds_mn + kk * 128), umma_desc(k_mn + kk * 128), idesc_dq,
                  kk > 0 ? 1u : 0u);
      };

      mbar_wait(&bars[kBarKV], 0);
      if (warp == 13) {
        // ---------------- stream X: S^T and dV ----------------
        mbar_wait(&bars[kBarQFull0], 0);
        tc_fence_after();
        // Q(it) is read by S^T(it) [X] and dK(it) [Y]: each stream releases it once its own reader is issued.
        if (elect_one()) {
          mma_kmajor(kColST, k_km, q_km);
          tc_commit(&bars[kBarSFull]);
          tc_commit(&bars[kBarQEmpty0]);
        }
        __syncwarp();
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t st = it & 1;
          mbar_wait(&bars[kBarDOFull0 + st], (it >> 1) & 1);
          mbar_wait(&bars[kBarPReady], it & 1);
          // dP^T(it) heads the loop that sets the iteration period (dP -> dS -> dK.dQ -> drain -> dP(next)): let stream Y
          // put it into the tensor pipe first; dV(it) and S^T(it+1) have a whole iteration of slack
          mbar_wait(&bars[kBarDPIssued], it & 1);
          tc_fence_after();
          if (lane == 0) FA_TRACE(0, it);
          if (elect_one()) {
            mma_from_tmem(kColDV, kColST, do_mn + st * kTileLo, it > 0);  // dV(it) += P^T dO
            tc_commit(&bars[kBarDOEmpty0 + st]);
          }
          __syncwarp();
          if (it + 1 < n_iter) {
            mbar_wait(&bars[kBarQFull0 + (st ^ 1)], ((it + 1) >> 1) & 1);
            tc_fence_after();
            if (lane == 0) FA_TRACE(1, it);
            if (elect_one()) {
              // S^T(it+1): P^T(it) in the same columns has been consumed by dV(it) (in-order within this stream)
              mma_kmajor(kColST, k_km, q_km + (st ^ 1) * kTileLo);
              tc_commit(&bars[kBarSFull]);
              tc_commit(&bars[kBarQEmpty0 + (st ^ 1)]);
            }
            __syncwarp();
          }
        }
      } else {
        // ---------------- stream Y: dP^T, dK, dQ ----------------
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t st = it & 1;
          mbar_wait(&bars[kBarDOFull0 + st], (it >> 1) & 1);
          if (it > 0) mbar_wait(&bars[kBarDQDrained], (it - 1) & 1);  // dP^T reuses the dQ(it-1) columns
          tc_fence_after();
          if (lane == 0) FA_TRACE(2, it);
          if (elect_one()) {
            mma_kmajor(kColDPT, v_km, do_km + st * kTileLo);  // dP^T(it) = V dO^T
            tc_commit(&bars[kBarDPFull]);
            tc_commit(&bars[kBarDOEmpty0 + st]);
            mbar_arrive(&bars[kBarDPIssued]);
          }
          __syncwarp();
          mbar_wait(&bars[kBarQFull0 + st], (it >> 1) & 1);
          // dK(it) += dS^T Q reads the packed dS^T straight from the DPT columns (A operand in TMEM): its first half
          // starts as soon as the first 32 queries of each warpgroup are stored ...
          mbar_wait(&bars[kBarDSHalf], it & 1);
          tc_fence_after();
          if (elect_one()) mma_from_tmem_part(kColDK, kColDPT, q_mn + st * kTileLo, it > 0, 0);
          __syncwarp();
          mbar_wait(&bars[kBarDSReady], it & 1);
          tc_fence_after();
          if (lane == 0) FA_TRACE(3, it);
          if (elect_one()) {
            mma_from_tmem_part(kColDK, kColDPT, q_mn + st * kTileLo, true, 1);  // ... the second half once all are
            tc_commit(&bars[kBarQEmpty0 + st]);
            mma_dq();  // dQ(it) = dS K overwrites the DPT columns: it must follow dK(it) in this (in-order) stream
            tc_commit(&bars[kBarDQFull]);
          }
          __syncwarp();
        }
      }
      tc_commit_elect(&bars[kBarDKVDone]);
    }
    __syncwarp();
  } else if (warp >= 8 && warp < 12) {
    // ===================================== dQ drain warpgroup =====================================
    const int row = threadIdx.x - 256;  // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    // dQ(it) sits in the DPT columns.  Two 64-column halves: each is read into registers, written to two 16 KiB
    // staging buffers (128-byte swizzled rows of 32 fp32) and reduce-added into dq_accum by TMA.  The staging buffers
    // are the dO stage of this very tile: dO(it) is dead once dV(it) and dP^T(it) have run, and dO(it+2) is not needed
    // for more than a full iteration, so the slow L2 reduce never sits on anyone's critical path.
    // `dq_drained` is signalled as soon as the LAST TMEM read has landed; `stage_free` (which gates the producer's
    // next load into this dO stage) once the last TMA read of the staging buffers has finished.
    for (int it = 0, i = first_tile; it < n_iter; ++it, i = next_tile(i)) {
      const int st = it & 1;
      uint8_t* dq_smem = Cfg::kStageInDO ? do_smem + st * Cfg::kTileBytes : smem + Cfg::kOffStage;
      mbar_wait(&bars[kBarDQFull], it & 1);
      if constexpr (Cfg::kStageInDO)
        mbar_wait(&bars[kBarDOEmpty0 + st], (it >> 1) & 1);  // dV(it) (other MMA stream) has finished reading dO(it)
      tc_fence_after();
      if (row == 0) FA_TRACE(8, it);
      float v[64];
      // 32 fp32 columns of this row -> one 128-byte swizzled row of staging buffer `cb`
      auto stage_chunk = [&](int cb, int off) {
        uint8_t* rowp = dq_smem + cb * Cfg::kDqStageBytes + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(rowp + ((c ^ (row & 7)) << 4)) =
              make_float4(v[off + 4 * c], v[off + 4 * c + 1], v[off + 4 * c + 2], v[off + 4 * c + 3]);
      };
      auto reduce_chunk = [&](int cb, int col) {
        if (col < p.d) tma_reduce_add_3d(&tm_dq, dq_smem + cb * Cfg::kDqStageBytes, col, i * kT, bh);
        tma_store_commit();
      };
      tmem_ld32(tmem_base + lane_sel + kColDPT, reinterpret_cast<uint32_t*>(v));
      tmem_ld32(tmem_base + lane_sel + kColDPT + 32, reinterpret_cast<uint32_t*>(v) + 32);
      tc_wait_ld();
      stage_chunk(0, 0);
      stage_chunk(1, 32);
      if (D == 128) {  // second 64 columns straight away: TMEM goes back before any TMA bookkeeping
        tmem_ld32(tmem_base + lane_sel + kColDPT + 64, reinterpret_cast<uint32_t*>(v));
        tmem_ld32(tmem_base + lane_sel + kColDPT + 96, reinterpret_cast<uint32_t*>(v) + 32);
        tc_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&bars[kBarDQDrained]);
      if (row == 0) FA_TRACE(9, it);
      fence_proxy_async_smem();
      named_bar_sync(3, 128);
      if (row == 0) {
        reduce_chunk(0, 0);
        reduce_chunk(1, 32);
      }
      if (D == 128) {
        // columns 64-127 follow through the same two buffers, each as soon as its previous reduce has been read, so
        // the reduce engine (about 40 B/ns per SM, measured) never waits for the staging stores of a whole half
        if (row == 0) {
          tma_store_wait_read<1>();
          FA_TRACE(10, it);
        }
        named_bar_sync(3, 128);
        stage_chunk(0, 0);
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (row == 0) {
          reduce_chunk(0, 64);
          tma_store_wait_read<1>();
        }
        named_bar_sync(3, 128);
        stage_chunk(1, 32);
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (row == 0) reduce_chunk(1, 96);
      }
      if (row == 0) {
        FA_TRACE(14, it);
        tma_store_wait_read<0>();
        mbar_arrive(&bars[kBarStageFree0 + st]);
        FA_TRACE(11, it);
      }
      if constexpr (!Cfg::kStageInDO) named_bar_sync(3, 128);  // dedicated staging is reused by the very next tile
    }
    if (row == 0) tma_store_wait_exit();  // staging read; the reduce itself completes by grid end
  } else if (warp < 8) {
    // ===================================== compute warpgroups =====================================
    const int wg = warp >> 2;
    const int r = threadIdx.x & 127;  // kv row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int col_base = wg * 64;
    const uint32_t t_st = tmem_base + lane_sel + kColST + col_base;
    const uint32_t t_dpt = tmem_base + lane_sel + kColDPT + col_base;
    uint8_t* ds_row = ds_smem + wg * (kT * 128) + r * 128;

    const bool dropout = kExt && p.drop_threshold != 0;
    for (int it = 0, i = first_tile; it < n_iter; ++it, i = next_tile(i)) {
      const int st = it & 1;
      const float* ls = stats_smem + st * 256 + col_base;  // -lse * log2e for this WG's 64 query columns
      const float* dl = ls + 128;                          // -delta
      // key (j*128 + r) is visible to query (i*128 + c) iff c >= c_min
      // ... and keys past n_kv (zero-filled by TMA) are never visible
      const bool kv_tail = (j * kT + kT > p.n_kv);
      int c_min = p.causal ? r + (j - i) * kT - p.diag : 0;
      if (j * kT + r >= p.n_kv) c_min = 1 << 30;
      const bool need_mask = kv_tail || (p.causal && (kT - 1 + (j - i) * kT - p.diag > 0));

      mbar_wait(&bars[kBarQFull0 + st], (it >> 1) & 1);  // row statistics ride on the Q barrier
      mbar_wait(&bars[kBarSFull], it & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(4, it);
      if (threadIdx.x == 0 && it == 0) FA_LIFE(3, clock64());  // first scores have arrived
#ifdef FA_BWD_TRACE
      if (threadIdx.x == 0 && fa_bwd_lin_block() == fa_bwd_trace_block && it < FA_BWD_TRACE_ITERS) {
        long long gt;  // wall-clock ns next to the cycle stamp: calibrates the SM clock under load
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        fa_bwd_trace_buf[15 * FA_BWD_TRACE_ITERS + it] = gt;
      }
#endif
      float pr[64];
      uint32_t keep[2] = {0xffffffffu, 0xffffffffu};  // kExt + dropout: keep bit of each of this thread's 64 queries
      const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s[32];
        tmem_ld32(t_st + c * 32, reinterpret_cast<uint32_t*>(s));
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int x4 = 0; x4 < 8; ++x4) {
          const float4 l4 = *reinterpret_cast<const float4*>(ls + c * 32 + x4 * 4);  // -lse * log2e
          const float2 t0 = ffma2(make_float2(s[x4 * 4], s[x4 * 4 + 1]), c2, make_float2(l4.x, l4.y));
          const float2 t1 = ffma2(make_float2(s[x4 * 4 + 2], s[x4 * 4 + 3]), c2, make_float2(l4.z, l4.w));
          float pv[4] = {ex2(t0.x), ex2(t0.y), ex2(t1.x), ex2(t1.y)};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int x = x4 * 4 + e;
            if (need_mask) pv[e] = (col_base + c * 32 + x >= c_min) ? pv[e] : 0.f;
            pr[c * 32 + x] = pv[e];
          }
        }
        if constexpr (kExt) {
          if (dropout) {
            // One Philox call covers 4 queries (words) x 4 keys (bytes) and the four lanes of a quad are four consecutive
            // keys (same key >> 2), so they need the SAME calls, each lane its own byte (key & 3) of every word: lane L
            // computes the calls of query groups x4 with (x4 & 3) == (L & 3) and the words go round the quad with
            // xor-shuffles -- 2 calls + 24 shuffles per 32 queries instead of 8 calls.
            const uint32_t kg = static_cast<uint32_t>(p.kv_col0 + j * kT + r);
            const uint32_t qg0 = static_cast<uint32_t>(p.q_row0 + i * kT + col_base + c * 32);
            const uint32_t me = lane & 3, shift = (kg & 3) * 8;
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
              const Philox4 w = philox4x32_7((qg0 >> 2) + 4 * jq + me, kg >> 2, static_cast<uint32_t>(bh), p.rng_offset,
                                             p.seed_lo, p.seed_hi);
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) {  // after this sub-round: the call of query group 4 jq + (me ^ kx)
                const uint32_t x4 = 4 * jq + (me ^ kx);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t word = kx == 0 ? w.w[e] : __shfl_xor_sync(0xffffffffu, w.w[e], kx);
                  if (((word >> shift) & 0xFFu) < p.drop_threshold) keep[c] &= ~(1u << (x4 * 4 + e));
                }
              }
            }
          }
        }
        // P^T as the dV operand carries the dropout (P o keep / (1 - p)); pr keeps the un-dropped P for dS
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float a = pr[c * 32 + 2 * x], b = pr[c * 32 + 2 * x + 1];
          if constexpr (kExt) {
            if (dropout) {
              a = ((keep[c] >> (2 * x)) & 1u) ? a * p.drop_scale : 0.f;
              b = ((keep[c] >> (2 * x + 1)) & 1u) ? b * p.drop_scale : 0.f;
            }
          }
          pk[x] = pack2<kBF16>(a, b);
        }
        tmem_st16(t_st + c * 16, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&bars[kBarPReady]);
      if (threadIdx.x == 0) FA_TRACE(5, it);
      mbar_wait(&bars[kBarDPFull], it & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(6, it);
      {
        // dS^T = P o (dP^T - delta) in four 16-column steps; the TMEM load of step c+1 is in flight while step c is
        // computed, converted and written twice: packed over the consumed dP^T columns in TMEM (A operand of dK) and to
        // the swizzled shared-memory tile (A operand of dQ, which needs the un-transposed view)
        float dp[2][16];
        tmem_ld16(t_dpt, reinterpret_cast<uint32_t*>(dp[0]));
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c + 1 < 4) tmem_ld16(t_dpt + (c + 1) * 16, reinterpret_cast<uint32_t*>(dp[(c + 1) & 1]));
          float* dpc = dp[c & 1];
          if constexpr (kExt) {
            if (dropout) {  // dP = (dO V^T) o keep / (1 - p)
#pragma unroll
              for (int x = 0; x < 16; ++x)
                dpc[x] = ((keep[c >> 1] >> ((c & 1) * 16 + x)) & 1u) ? dpc[x] * p.drop_scale : 0.f;
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int x4 = 0; x4 < 4; ++x4) {
            const float4 d4 = *reinterpret_cast<const float4*>(dl + c * 16 + x4 * 4);  // -delta
            const int x = x4 * 4;
            const float2 a = fmul2(make_float2(pr[c * 16 + x], pr[c * 16 + x + 1]),
                                   fadd2(make_float2(dpc[x], dpc[x + 1]), make_float2(d4.x, d4.y)));
            const float2 b = fmul2(make_float2(pr[c * 16 + x + 2], pr[c * 16 + x + 3]),
                                   fadd2(make_float2(dpc[x + 2], dpc[x + 3]), make_float2(d4.z, d4.w)));
            pk[x >> 1] = pack2<kBF16>(a.x, a.y);
            pk[(x >> 1) + 1] = pack2<kBF16>(b.x, b.y);
          }
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const int chunk = c * 2 + ch;
            *reinterpret_cast<uint4*>(ds_row + ((chunk ^ (r & 7)) << 4)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          }
          if (c + 1 < 4) tc_wait_ld();
          tmem_st8(t_dpt + c * 8, pk);  // packed dS^T over the dP^T columns this thread has already consumed
          if (c == 1) {  // first 32 queries of this warpgroup are in TMEM: dK can start on them
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&bars[kBarDSHalf]);
          }
        }
      }
      tc_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bars[kBarDSReady]);
      if (threadIdx.x == 0) FA_TRACE(7, it);
    }

    // ------------------------------- epilogue: WG0 stores dV, WG1 stores dK * scale -------------------------------
    if (threadIdx.x == 0) FA_LIFE(4, clock64());  // last dS handed over
    if (n_iter > 0) {
      mbar_wait(&bars[kBarDKVDone], 0);
      tc_fence_after();
      if (threadIdx.x == 0) FA_LIFE(5, clock64());  // last dK/dV products finished
    } else {
      mbar_wait(&bars[kBarKV], 0);  // staging reuses the K/V tiles: their loads must have landed
    }
    uint8_t* stage_tile = wg == 0 ? v_smem : k_smem;
    const uint32_t t_acc = tmem_base + lane_sel + (wg == 0 ? kColDV : kColDK);
    const float mul = wg == 0 ? 1.f : p.scale;
    const CUtensorMap* tm = wg == 0 ? &tm_dv : &tm_dk;
    if (!p.accum_kv) {
      // 16-bit results: the whole tile goes through the (dead) V / K buffer and out with one TMA store per 64 columns
#pragma unroll
      for (int q4 = 0; q4 < D / 32; ++q4) {
        float a[32];
        if (n_iter > 0) {
          tmem_ld32(t_acc + q4 * 32, reinterpret_cast<uint32_t*>(a));
          tc_wait_ld();
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) a[x] = 0.f;
        }
        uint32_t pk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) pk[x] = pack2<kBF16>(a[2 * x] * mul, a[2 * x + 1] * mul);
        uint8_t* sub = stage_tile + (q4 >> 1) * kSub + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (q4 & 1) * 4 + ch;
          *reinterpret_cast<uint4*>(sub + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + wg, 128);
      if (r == 0) {
        for (int ch = 0; ch < kChunks; ++ch) tma_store_3d(tm, stage_tile + ch * kSub, ch * 64, j * kT, bh);
        tma_store_commit();
        tma_store_wait_exit();  // K/V staging has been read; the stores complete by grid end
      }
    } else if (n_iter > 0 || p.accum_kv == 2) {
      // fp32 partials for ring attention (tm_dv / tm_dk describe fp32 tensors here): reduce-added into the given
      // accumulators (mode 1) or stored (mode 2, zeros where this K/V tile saw no query), 32 columns -- one 128-byte
      // swizzled row per kv row, 16 KiB -- at a time, D/64 such chunks per round through the same V / K buffer.
      constexpr int kPerRound = D / 64;
#pragma unroll
      for (int round = 0; round < 2; ++round) {
#pragma unroll
        for (int u = 0; u < kPerRound; ++u) {
          float a[32];
          if (n_iter > 0) {
            tmem_ld32(t_acc + (round * kPerRound + u) * 32, reinterpret_cast<uint32_t*>(a));
            tc_wait_ld();
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x) a[x] = 0.f;
          }
          uint8_t* rowp = stage_tile + u * Cfg::kDqStageBytes + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(rowp + ((c ^ (r & 7)) << 4)) =
                make_float4(a[4 * c] * mul, a[4 * c + 1] * mul, a[4 * c + 2] * mul, a[4 * c + 3] * mul);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + wg, 128);
        if (r == 0) {
          for (int u = 0; u < kPerRound; ++u) {
            const int col = (round * kPerRound + u) * 32;
            if (col >= p.d) continue;
            if (p.accum_kv == 2) tma_store_3d(tm, stage_tile + u * Cfg::kDqStageBytes, col, j * kT, bh);
            else tma_reduce_add_3d(tm, stage_tile + u * Cfg::kDqStageBytes, col, j * kT, bh);
          }
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        named_bar_sync(1 + wg, 128);  // the buffer is free for the next round / safe to leave
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem_base, 512);
#ifdef FA_BWD_TRACE
  if (threadIdx.x == 0) {
    FA_LIFE(6, clock64());
    FA_LIFE(7, fa_bwd_globaltimer());
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// pre-pass: delta = rowsum(dO o O), packed with lse*log2e per 128-row tile (HBM-bound: 2*D*2 + 4 B read, 8 B written
// per query row).  One warp handles 32/(D/8) rows at a time with 16-byte loads.
// ------------------------------------------------------------------------------------------------
template <int DP, bool kBF16>
__global__ void __launch_bounds__(256)
fa_bwd_prepare_kernel(const uint16_t* __restrict__ o, const uint16_t* __restrict__ d_o, const float* __restrict__ lse,
                      float* __restrict__ rowstats, long long n_q, long long bh, long long nqt, long long q_bh_stride,
                      long long lse_bh_stride, int d, float* __restrict__ dq_zero) {
  constexpr int kLanesPerRow = DP / 8;  // DP = d rounded up to 64 / 128; lanes past d / 8 idle
  constexpr int kRowsPerWarp = 32 / kLanesPerRow;
  const int lane = threadIdx.x & 31;
  const int sub = lane / kLanesPerRow, li = lane % kLanesPerRow;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long rows_pad = nqt * kT;
  const long long total = bh * rows_pad;
  for (long long r0 = warp_global * kRowsPerWarp; r0 < total; r0 += nwarps * kRowsPerWarp) {
    const long long r = r0 + sub;  // padded row id (rows_pad is a multiple of kRowsPerWarp, so r < total)
    const long long b = r / rows_pad, rr = r % rows_pad;
    const bool valid = rr < n_q;
    float acc = 0.f;
    if (valid && li * 8 < d) {
      const long long off = b * q_bh_stride + rr * d + li * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + off);
      const uint4 g = *reinterpret_cast<const uint4*>(d_o + off);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack2<kBF16>(aw[e]), y = unpack2<kBF16>(gw[e]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
#pragma unroll
    for (int s = kLanesPerRow / 2; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (li == 0) {
      float nl2 = -INFINITY;  // padded rows and rows that saw no key: P = 2^(x - inf) = 0
      if (valid) {
        const float l = lse[b * lse_bh_stride + rr];
        if (l != -INFINITY) nl2 = -l * 1.4426950408889634f;
      }
      float* tile = rowstats + (b * nqt + rr / kT) * 256;
      tile[rr % kT] = nl2;                      // stored negated: the main kernel folds them in with one FMA / ADD
      tile[128 + rr % kT] = valid ? -acc : 0.f;
    }
  }
  // optional: zero-fill the fp32 dQ accumulator the main pass reduce-adds into (saves the caller a separate memset
  // launch; the stores overlap the O / dO reads above)
  if (dq_zero != nullptr) {
    const long long per_slice = n_q * d / 4;  // float4 per slice (d % 8 == 0)
    const long long total_vec = bh * per_slice;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
      const long long b = i / per_slice, e = i % per_slice;
      reinterpret_cast<float4*>(dq_zero + b * q_bh_stride)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int D, bool kBF16, bool kExt>
static int launch_bwd(const Geometry& g, const ExtArgs& ext, const void* q, const void* k, const void* v,
                      const void* d_o, const float* rowstats, float* dq_accum, void* dk, void* dv, int accum_kv,
                      long long acc_bh_stride, cudaStream_t stream) {
  using Cfg = BwdCfg<D>;
  constexpr int kSmem = kExt ? Cfg::kSmemBytesExt : Cfg::kSmemBytes;
  const int elem = kBF16 ? kElemBF16 : kElemF16;
  CUtensorMap tm_q, tm_k, tm_v, tm_do, tm_dq, tm_dk, tm_dv;
  int rc;
  if ((rc = make_tmap_3d(&tm_q, q, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kT))) return rc;
  if ((rc = make_tmap_3d(&tm_do, d_o, elem, g.d, g.n_q, g.bh, g.q_bh_stride, 64, kT))) return rc;
  if ((rc = make_tmap_3d(&tm_k, k, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kT))) return rc;
  if ((rc = make_tmap_3d(&tm_v, v, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kT))) return rc;
  if (accum_kv) {  // fp32 accumulators, 32-column boxes like dq_accum
    if ((rc = make_tmap_3d(&tm_dk, dk, kElemF32, g.d, g.n_kv, g.bh, acc_bh_stride, 32, kT))) return rc;
    if ((rc = make_tmap_3d(&tm_dv, dv, kElemF32, g.d, g.n_kv, g.bh, acc_bh_stride, 32, kT))) return rc;
  } else {
    if ((rc = make_tmap_3d(&tm_dk, dk, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kT))) return rc;
    if ((rc = make_tmap_3d(&tm_dv, dv, elem, g.d, g.n_kv, g.bh, g.kv_bh_stride, 64, kT))) return rc;
  }
  if ((rc = make_tmap_3d(&tm_dq, dq_accum, kElemF32, g.d, g.n_q, g.bh, g.q_bh_stride, 32, kT))) return rc;

  BwdParams p;
  p.rowstats = rowstats;
  p.dq_accum = dq_accum;
  p.dq_bh_stride = g.q_bh_stride;
  p.n_q = static_cast<int>(g.n_q);
  p.n_kv = static_cast<int>(g.n_kv);
  p.bh = static_cast<int>(g.bh);
  p.causal = g.causal;
  p.diag = g.diag;
  p.d = g.d;
  p.accum_kv = accum_kv;
  p.nqt = static_cast<int>((g.n_q + kT - 1) / kT);
  p.nkt = static_cast<int>((g.n_kv + kT - 1) / kT);
  p.group_log2 = sched_group_log2(g.causal != 0, p.nkt, g.bh);
  while (((g.bh + (1ll << p.group_log2) - 1) >> p.group_log2) > 65535) ++p.group_log2;  // grid.z limit
  p.scale = g.scale;
  p.scale_log2 = g.scale * 1.4426950408889634f;

  p.block_mask = ext.block_mask;
  p.mask_bh_stride = ext.mask_bh_stride;
  p.seed_lo = ext.seed_lo;
  p.seed_hi = ext.seed_hi;
  p.rng_offset = ext.rng_offset;
  p.drop_threshold = ext.drop_threshold;
  p.drop_scale = ext.drop_scale;
  p.q_row0 = ext.q_row0;
  p.kv_col0 = ext.kv_col0;
  auto kern = fa_bwd_kernel<D, kBF16, kExt>;
  static bool attr_set[64];  // per device (function attributes are per context); benign race: idempotent
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
      return FA_SM100_ELAUNCH;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long rank_lo = p.nkt < (1 << kRankBitsY) ? p.nkt : (1 << kRankBitsY);
  const long long rank_hi = (p.nkt + (1 << kRankBitsY) - 1) >> kRankBitsY;
  const long long gx = rank_hi << p.group_log2, gz = (g.bh + (1ll << p.group_log2) - 1) >> p.group_log2;
  if (gx > 0x7fffffffll || gz > 65535) return FA_SM100_EINVAL_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rank_lo), static_cast<unsigned>(gz));
  kern<<<grid, kBwdThreads, kSmem, stream>>>(tm_q, tm_k, tm_v, tm_do, tm_dq, tm_dk, tm_dv, p);
  return launch_status();
}

}  // namespace fa

extern "C" size_t fa_sm100_rowstats_bytes(const fa_sm100_shape* s) {
  if (!s || s->bh <= 0 || s->n_q <= 0) return 0;
  const size_t nqt = static_cast<size_t>((s->n_q + fa::kT - 1) / fa::kT);
  return static_cast<size_t>(s->bh) * nqt * 256 * sizeof(float);
}

extern "C" int fa_sm100_bwd_prepare(const fa_sm100_shape* s, const void* o, const void* d_o, const float* lse,
                                    float* rowstats, float* dq_accum_zero, void* stream) {
  fa::Geometry g;
  int rc = fa::check_shape(s, &g, /*max_d=*/256);
  if (rc) return rc;
  if (!fa::aligned16(o) || !fa::aligned16(d_o) || lse == nullptr || !fa::aligned16(rowstats))
    return FA_SM100_EINVAL_PTR;
  if (dq_accum_zero != nullptr && !fa::aligned16(dq_accum_zero)) return FA_SM100_EINVAL_PTR;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nqt = (g.n_q + fa::kT - 1) / fa::kT;
  const long long rows = g.bh * nqt * fa::kT;
  const int rows_per_block = (256 / 32) * (32 / (g.dp / 8));
  long long grid = (rows + rows_per_block - 1) / rows_per_block;
  if (grid > 148 * 16) grid = 148 * 16;
  const uint16_t* op = static_cast<const uint16_t*>(o);
  const uint16_t* gp = static_cast<const uint16_t*>(d_o);
  const bool bf = g.dtype == FA_SM100_DTYPE_BF16;
#define FA_LAUNCH_PREP(DD, BF)                                                                                  \
  fa::fa_bwd_prepare_kernel<DD, BF><<<static_cast<unsigned>(grid), 256, 0, st>>>(op, gp, lse, rowstats, g.n_q, g.bh, \
                                                                                  nqt, g.q_bh_stride, g.lse_bh_stride, g.d, dq_accum_zero)
  if (g.dp == 256) {
    if (bf) FA_LAUNCH_PREP(256, true); else FA_LAUNCH_PREP(256, false);
  } else if (g.dp == 128) {
    if (bf) FA_LAUNCH_PREP(128, true); else FA_LAUNCH_PREP(128, false);
  } else {
    if (bf) FA_LAUNCH_PREP(64, true); else FA_LAUNCH_PREP(64, false);
  }
#undef FA_LAUNCH_PREP
  return fa::launch_status();
}

namespace fa {
// head dims 129..256: csrc/fa_bwd256_sm100.cu
int bwd256_dispatch(const Geometry& g, const void* q, const void* k, const void* v, const void* d_o,
                    const float* rowstats, float* dq_accum, void* dk, void* dv, cudaStream_t st);

static int bwd_dispatch(const fa_sm100_shape* s, const fa_sm100_ext* ext, const void* q, const void* k, const void* v,
                        const void* d_o, const float* rowstats, float* dq_accum, void* dk, void* dv, int accum_kv,
                        long long acc_bh_stride, void* stream) {
  Geometry g;
  int rc = check_shape(s, &g, /*max_d=*/256);
  if (rc) return rc;
  ExtArgs ea;
  if ((rc = check_ext(ext, s, kMaxMaskTiles, &ea))) return rc;
  // the ring-accumulator and block-sparse / dropout forms stop at head dim 128
  if (g.dp == 256 && (accum_kv || ea.block_mask != nullptr || ea.drop_threshold != 0)) return FA_SM100_EINVAL_HEADDIM;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(d_o) || !aligned16(rowstats) ||
      !aligned16(dq_accum) || !aligned16(dk) || !aligned16(dv))
    return FA_SM100_EINVAL_PTR;
  if (accum_kv) {
    if (acc_bh_stride == 0) acc_bh_stride = g.n_kv * g.d;
    if (acc_bh_stride < g.n_kv * g.d || (acc_bh_stride % 4)) return FA_SM100_EINVAL_SHAPE;
  }
  if ((rc = check_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g.dp == 256) return bwd256_dispatch(g, q, k, v, d_o, rowstats, dq_accum, dk, dv, st);
  const bool bf = g.dtype == FA_SM100_DTYPE_BF16;
#define FA_BWD_GO(DD, BF, EXT) \
  launch_bwd<DD, BF, EXT>(g, ea, q, k, v, d_o, rowstats, dq_accum, dk, dv, accum_kv, acc_bh_stride, st)
  if (ea.block_mask == nullptr && ea.drop_threshold == 0) {  // the dense kernel
    if (g.dp == 128) return bf ? FA_BWD_GO(128, true, false) : FA_BWD_GO(128, false, false);
    return bf ? FA_BWD_GO(64, true, false) : FA_BWD_GO(64, false, false);
  }
  if (g.dp == 128) return bf ? FA_BWD_GO(128, true, true) : FA_BWD_GO(128, false, true);
  return bf ? FA_BWD_GO(64, true, true) : FA_BWD_GO(64, false, true);
#undef FA_BWD_GO
}
}  // namespace fa

extern "C" int fa_sm100_bwd(const fa_sm100_shape* s, const void* q, const void* k, const void* v, const void* d_o,
                            const float* rowstats, float* dq_accum, void* dk, void* dv, void* stream) {
  return fa::bwd_dispatch(s, nullptr, q, k, v, d_o, rowstats, dq_accum, dk, dv, 0, 0, stream);
}

extern "C" int fa_sm100_bwd_ex(const fa_sm100_shape* s, const fa_sm100_ext* ext, const void* q, const void* k,
                               const void* v, const void* d_o, const float* rowstats, float* dq_accum, void* dk,
                               void* dv, void* stream) {
  return fa::bwd_dispatch(s, ext, q, k, v, d_o, rowstats, dq_accum, dk, dv, 0, 0, stream);
}

extern "C" int fa_sm100_bwd_accum(const fa_sm100_shape* s, const void* q, const void* k, const void* v,
                                  const void* d_o, const float* rowstats, float* dq_accum, float* dk_accum,
                                  float* dv_accum, int64_t acc_bh_stride, int32_t overwrite, void* stream) {
  return fa::bwd_dispatch(s, nullptr, q, k, v, d_o, rowstats, dq_accum, dk_accum, dv_accum, overwrite ? 2 : 1,
                          acc_bh_stride, stream);
}

#ifdef FA_BWD_TRACE
// debug build only: select the CTA to trace / copy its timeline (events x FA_BWD_TRACE_ITERS clock64 values) to the host
extern "C" int fa_sm100_debug_bwd_life(long long* host_dst, int n_ctas) {
  if (!host_dst || n_ctas <= 0 || n_ctas > FA_BWD_LIFE_MAX_CTAS) return -1;
  return cudaMemcpyFromSymbol(host_dst, fa_bwd_life_buf, sizeof(long long) * 8 * n_ctas) == cudaSuccess ? n_ctas : -1;
}
extern "C" int fa_sm100_debug_bwd_trace(int set_block, long long* host_dst, int max_values) {
  if (set_block >= 0) {
    long long zero[FA_BWD_TRACE_EVENTS * FA_BWD_TRACE_ITERS] = {};
    cudaMemcpyToSymbol(fa_bwd_trace_buf, zero, sizeof(zero));
    return cudaMemcpyToSymbol(fa_bwd_trace_block, &set_block, sizeof(int)) == cudaSuccess ? 0 : -1;
  }
  const int n = FA_BWD_TRACE_EVENTS * FA_BWD_TRACE_ITERS;
  if (!host_dst || max_values < n) return -1;
  return cudaMemcpyFromSymbol(host_dst, fa_bwd_trace_buf, n * sizeof(long long)) == cudaSuccess ? n : -1;
}
#endif
