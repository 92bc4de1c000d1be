// fp32 attention forward + backward for fp32 INPUTS (BASELINE config C1: B2 H4 N512 d64 fp32; reference
// csrc/fa1/fa1_fwd.cu:67,79-80 up-casts every input to fp32 and computes in fp32, tests/utils.py:31-36 holds fp32 to
// rtol = atol = 1e-4).  Meeting 1e-4 on the tensor cores would need error-compensated splits of every operand -- Q, K,
// V, P, dS -- (kind::tf32 alone leaves ~5e-4); this path instead does the arithmetic the reference does: plain fp32 FMA
// on the CUDA cores, tiled through shared memory, same online-softmax recurrence (src/fa1/torch/impl.py:26-68) and
// KV-outer backward (:70-115).  It is HBM/FMA-bound SIMT code, roughly 1/40 of the tcgen05 path's rate, and exists so
// that fp32 callers of the reference keep working on the GPU with the reference's own tolerances -- the 16-bit path is
// the product's fast path.
#include "ptx.cuh"
#include "fa_host.cuh"

namespace fa {

constexpr int kF32Tile = 64;  // query rows and key rows per tile

struct F32Params {
  long long q_bh_stride, kv_bh_stride, lse_bh_stride;
  int n_q, n_kv, d, causal, diag;
  float scale;
};

// rows [row0, row0 + 64) x d of a dense-row matrix -> shared [64][DP], zero beyond the end (16-byte loads, d % 4 == 0)
template <int DP, int kThreads>
__device__ __forceinline__ void f32_load_tile(float* dst, const float* src, int row0, int n_rows, int d) {
  const int vec_per_row = d >> 2;
  for (int idx = threadIdx.x; idx < kF32Tile * vec_per_row; idx += kThreads) {
    const int r = idx / vec_per_row, c = (idx - r * vec_per_row) << 2;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_rows) x = *reinterpret_cast<const float4*>(src + static_cast<long long>(row0 + r) * d + c);
    *reinterpret_cast<float4*>(dst + r * DP + c) = x;
  }
}

// ------------------------------------------------------------------------------------------------
// forward: CTA = 64 query rows of one slice, 128 threads as a 16 x 8 grid; thread (ty, tx) owns score rows 4 ty .. 4 ty + 3,
// score columns tx + 8 b (b < 8), and output columns 4 tx + 32 e .. + 3 (e < DMAX / 32) of the same rows.
// ------------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(128) fa_f32_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                         const float* __restrict__ v, float* __restrict__ o,
                                                         float* __restrict__ lse, const F32Params p) {
  constexpr int DP = DMAX + 4, PP = kF32Tile + 1, kCols = DMAX / 32;
  extern __shared__ __align__(16) float smem_f32[];
  float* sQ = smem_f32;
  float* sK = sQ + kF32Tile * DP;
  float* sV = sK + kF32Tile * DP;
  float* sP = sV + kF32Tile * DP;
  const int bh = blockIdx.y, row0 = blockIdx.x * kF32Tile;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  const float* qb = q + static_cast<long long>(bh) * p.q_bh_stride;
  const float* kb = k + static_cast<long long>(bh) * p.kv_bh_stride;
  const float* vb = v + static_cast<long long>(bh) * p.kv_bh_stride;

  f32_load_tile<DP, 128>(sQ, qb, row0, p.n_q, p.d);
  float m[4], l[4], acc[4][kCols * 4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m[a] = -INFINITY;
    l[a] = 0.f;
#pragma unroll
    for (int e = 0; e < kCols * 4; ++e) acc[a][e] = 0.f;
  }
  int n_tiles = (p.n_kv + kF32Tile - 1) / kF32Tile;
  if (p.causal) {  // tiles whose first key is hidden from the tile's last query are skipped (src/fa1/torch/impl.py:15-16)
    const long long last_vis = static_cast<long long>(row0) + kF32Tile - 1 + p.diag;
    const int lim = last_vis < 0 ? 0 : static_cast<int>(last_vis / kF32Tile) + 1;
    n_tiles = lim < n_tiles ? lim : n_tiles;
  }
  for (int t = 0; t < n_tiles; ++t) {
    const int col0 = t * kF32Tile;
    __syncthreads();  // previous tile's K / V / P are no longer read
    f32_load_tile<DP, 128>(sK, kb, col0, p.n_kv, p.d);
    f32_load_tile<DP, 128>(sV, vb, col0, p.n_kv, p.d);
    __syncthreads();
    float s[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) s[a][b] = 0.f;
    for (int kk = 0; kk < p.d; kk += 4) {
      float4 qa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(sQ + (ty * 4 + a) * DP + kk);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const float4 kv4 = *reinterpret_cast<const float4*>(sK + (tx + 8 * b) * DP + kk);
#pragma unroll
        for (int a = 0; a < 4; ++a)
          s[a][b] = fmaf(qa[a].x, kv4.x, fmaf(qa[a].y, kv4.y, fmaf(qa[a].z, kv4.z, fmaf(qa[a].w, kv4.w, s[a][b]))));
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = row0 + ty * 4 + a;
      float mx = -INFINITY;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int c = col0 + tx + 8 * b;
        const bool vis = c < p.n_kv && (!p.causal || c <= r + p.diag);
        s[a][b] = vis ? s[a][b] * p.scale : -INFINITY;
        mx = fmaxf(mx, s[a][b]);
      }
#pragma unroll
      for (int sh = 1; sh < 8; sh <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
      const float m_new = fmaxf(m[a], mx);
      const float m_safe = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = expf(m[a] - m_safe);  // exp(-inf) = 0 on the first visible tile
      float sum = 0.f;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const float pv = expf(s[a][b] - m_safe);
        sum += pv;
        sP[(ty * 4 + a) * PP + tx + 8 * b] = pv;
      }
#pragma unroll
      for (int sh = 1; sh < 8; sh <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
      l[a] = l[a] * alpha + sum;
      m[a] = m_new;
#pragma unroll
      for (int e = 0; e < kCols * 4; ++e) acc[a][e] *= alpha;
    }
    __syncthreads();
    for (int kv = 0; kv < kF32Tile; ++kv) {
      float pa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pa[a] = sP[(ty * 4 + a) * PP + kv];
#pragma unroll
      for (int e = 0; e < kCols; ++e) {
        const float4 v4 = *reinterpret_cast<const float4*>(sV + kv * DP + tx * 4 + 32 * e);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          acc[a][e * 4 + 0] = fmaf(pa[a], v4.x, acc[a][e * 4 + 0]);
          acc[a][e * 4 + 1] = fmaf(pa[a], v4.y, acc[a][e * 4 + 1]);
          acc[a][e * 4 + 2] = fmaf(pa[a], v4.z, acc[a][e * 4 + 2]);
          acc[a][e * 4 + 3] = fmaf(pa[a], v4.w, acc[a][e * 4 + 3]);
        }
      }
    }
  }
  float* ob = o + static_cast<long long>(bh) * p.q_bh_stride;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = row0 + ty * 4 + a;
    if (r >= p.n_q) continue;
    const float inv = l[a] > 0.f ? 1.f / l[a] : 0.f;
#pragma unroll
    for (int e = 0; e < kCols; ++e) {
      const int c = tx * 4 + 32 * e;
      if (c < p.d)
        *reinterpret_cast<float4*>(ob + static_cast<long long>(r) * p.d + c) =
            make_float4(acc[a][e * 4] * inv, acc[a][e * 4 + 1] * inv, acc[a][e * 4 + 2] * inv, acc[a][e * 4 + 3] * inv);
    }
    if (tx == 0) lse[static_cast<long long>(bh) * p.lse_bh_stride + r] = l[a] > 0.f ? m[a] + logf(l[a]) : -INFINITY;
  }
}

// delta[r] = sum_c dO[r, c] * O[r, c]  (reference csrc/fa1/fa1_bwd.cu:57); one warp per row
__global__ void __launch_bounds__(256) fa_f32_delta_kernel(const float* __restrict__ o, const float* __restrict__ d_o,
                                                           float* __restrict__ delta, long long bh, long long n_q, int d,
                                                           long long q_bh_stride) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= bh * n_q) return;
  const long long b = warp / n_q, r = warp % n_q;
  const float* po = o + b * q_bh_stride + r * d;
  const float* pg = d_o + b * q_bh_stride + r * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(po[c], pg[c], acc);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
  if (lane == 0) delta[b * n_q + r] = acc;
}

__global__ void __launch_bounds__(256) fa_f32_zero_kernel(float* __restrict__ x, long long bh, long long slice_elems,
                                                          long long bh_stride) {
  const long long per = slice_elems >> 2, total = bh * per;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride)
    reinterpret_cast<float4*>(x + (i / per) * bh_stride)[i % per] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// backward, KV-outer: CTA = 64 key rows of one slice, 256 threads as a 16 x 16 grid.  Per query tile: S and dP micro
// tiles (4 x 4 per thread) -> P, dS through shared memory -> dV += P^T dO, dK += dS^T Q in registers (4 rows x DMAX/16
// columns per thread), dQ = dS K added to global memory with fp32 atomics (dq is zero-filled by the entry point).
// ------------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(256) fa_f32_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                         const float* __restrict__ v, const float* __restrict__ d_o,
                                                         const float* __restrict__ lse, const float* __restrict__ delta,
                                                         float* __restrict__ dq, float* __restrict__ dk,
                                                         float* __restrict__ dv, const F32Params p) {
  constexpr int DP = DMAX + 4, PP = kF32Tile + 1, kCols = DMAX / 64;
  extern __shared__ __align__(16) float smem_f32[];
  float* sK = smem_f32;
  float* sV = sK + kF32Tile * DP;
  float* sQ = sV + kF32Tile * DP;
  float* sG = sQ + kF32Tile * DP;  // dO
  float* sP = sG + kF32Tile * DP;
  float* sS = sP + kF32Tile * PP;  // dS
  float* sL = sS + kF32Tile * PP;  // lse (64) then delta (64)
  const int bh = blockIdx.y, col0 = blockIdx.x * kF32Tile;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* qb = q + static_cast<long long>(bh) * p.q_bh_stride;
  const float* gb = d_o + static_cast<long long>(bh) * p.q_bh_stride;
  const float* kb = k + static_cast<long long>(bh) * p.kv_bh_stride;
  const float* vb = v + static_cast<long long>(bh) * p.kv_bh_stride;
  float* dqb = dq + static_cast<long long>(bh) * p.q_bh_stride;

  f32_load_tile<DP, 256>(sK, kb, col0, p.n_kv, p.d);
  f32_load_tile<DP, 256>(sV, vb, col0, p.n_kv, p.d);
  float acc_k[4][kCols * 4], acc_v[4][kCols * 4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < kCols * 4; ++e) acc_k[a][e] = acc_v[a][e] = 0.f;

  const int n_qt = (p.n_q + kF32Tile - 1) / kF32Tile;
  int t_min = 0;
  if (p.causal) {  // query tiles whose last row cannot see this tile's first key are skipped (src/fa1/torch/impl.py:89)
    const int first = col0 - p.diag;
    t_min = first > 0 ? first / kF32Tile : 0;
  }
  for (int t = t_min; t < n_qt; ++t) {
    const int row0 = t * kF32Tile;
    __syncthreads();
    f32_load_tile<DP, 256>(sQ, qb, row0, p.n_q, p.d);
    f32_load_tile<DP, 256>(sG, gb, row0, p.n_q, p.d);
    if (threadIdx.x < 128) {
      const int r = row0 + (threadIdx.x & 63);
      const bool is_lse = threadIdx.x < 64;
      float x = is_lse ? INFINITY : 0.f;  // rows past the end: P = exp(s - inf) = 0
      if (r < p.n_q) {
        x = is_lse ? lse[static_cast<long long>(bh) * p.lse_bh_stride + r] : delta[static_cast<long long>(bh) * p.n_q + r];
        if (is_lse && x == -INFINITY) x = INFINITY;  // a row that saw no key contributes nothing
      }
      sL[threadIdx.x] = x;
    }
    __syncthreads();
    // S[q, kv] and dP[q, kv] for query rows 4 ty + a, key columns tx + 16 b
    float s[4][4], dp[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = dp[a][b] = 0.f;
    for (int kk = 0; kk < p.d; kk += 4) {
      float4 qa[4], ga[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        qa[a] = *reinterpret_cast<const float4*>(sQ + (ty * 4 + a) * DP + kk);
        ga[a] = *reinterpret_cast<const float4*>(sG + (ty * 4 + a) * DP + kk);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float4 k4 = *reinterpret_cast<const float4*>(sK + (tx + 16 * b) * DP + kk);
        const float4 v4 = *reinterpret_cast<const float4*>(sV + (tx + 16 * b) * DP + kk);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          s[a][b] = fmaf(qa[a].x, k4.x, fmaf(qa[a].y, k4.y, fmaf(qa[a].z, k4.z, fmaf(qa[a].w, k4.w, s[a][b]))));
          dp[a][b] = fmaf(ga[a].x, v4.x, fmaf(ga[a].y, v4.y, fmaf(ga[a].z, v4.z, fmaf(ga[a].w, v4.w, dp[a][b]))));
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int rl = ty * 4 + a, r = row0 + rl;
      const float l_r = sL[rl], d_r = sL[64 + rl];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int cl = tx + 16 * b, c = col0 + cl;
        const bool vis = c < p.n_kv && (!p.causal || c <= r + p.diag);
        const float pv = vis ? expf(s[a][b] * p.scale - l_r) : 0.f;
        sP[rl * PP + cl] = pv;
        sS[rl * PP + cl] = pv * (dp[a][b] - d_r);
      }
    }
    __syncthreads();
    // dV[kv, :] += P^T dO and dK[kv, :] += dS^T Q for key rows 4 ty + a, columns 4 tx + 64 e
    for (int r = 0; r < kF32Tile; ++r) {
      float pa[4], sa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        pa[a] = sP[r * PP + ty * 4 + a];
        sa[a] = sS[r * PP + ty * 4 + a];
      }
#pragma unroll
      for (int e = 0; e < kCols; ++e) {
        const float4 g4 = *reinterpret_cast<const float4*>(sG + r * DP + tx * 4 + 64 * e);
        const float4 q4 = *reinterpret_cast<const float4*>(sQ + r * DP + tx * 4 + 64 * e);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          acc_v[a][e * 4 + 0] = fmaf(pa[a], g4.x, acc_v[a][e * 4 + 0]);
          acc_v[a][e * 4 + 1] = fmaf(pa[a], g4.y, acc_v[a][e * 4 + 1]);
          acc_v[a][e * 4 + 2] = fmaf(pa[a], g4.z, acc_v[a][e * 4 + 2]);
          acc_v[a][e * 4 + 3] = fmaf(pa[a], g4.w, acc_v[a][e * 4 + 3]);
          acc_k[a][e * 4 + 0] = fmaf(sa[a], q4.x, acc_k[a][e * 4 + 0]);
          acc_k[a][e * 4 + 1] = fmaf(sa[a], q4.y, acc_k[a][e * 4 + 1]);
          acc_k[a][e * 4 + 2] = fmaf(sa[a], q4.z, acc_k[a][e * 4 + 2]);
          acc_k[a][e * 4 + 3] = fmaf(sa[a], q4.w, acc_k[a][e * 4 + 3]);
        }
      }
    }
    // dQ[q, :] += dS K * scale for query rows 4 ty + a, columns 4 tx + 64 e
    float acc_q[4][kCols * 4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < kCols * 4; ++e) acc_q[a][e] = 0.f;
    for (int c = 0; c < kF32Tile; ++c) {
      float sa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) sa[a] = sS[(ty * 4 + a) * PP + c];
#pragma unroll
      for (int e = 0; e < kCols; ++e) {
        const float4 k4 = *reinterpret_cast<const float4*>(sK + c * DP + tx * 4 + 64 * e);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          acc_q[a][e * 4 + 0] = fmaf(sa[a], k4.x, acc_q[a][e * 4 + 0]);
          acc_q[a][e * 4 + 1] = fmaf(sa[a], k4.y, acc_q[a][e * 4 + 1]);
          acc_q[a][e * 4 + 2] = fmaf(sa[a], k4.z, acc_q[a][e * 4 + 2]);
          acc_q[a][e * 4 + 3] = fmaf(sa[a], k4.w, acc_q[a][e * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = row0 + ty * 4 + a;
      if (r >= p.n_q) continue;
#pragma unroll
      for (int e = 0; e < kCols; ++e) {
        const int c = tx * 4 + 64 * e;
        if (c < p.d) {
          float* dst = dqb + static_cast<long long>(r) * p.d + c;
          red_add_v4(dst, acc_q[a][e * 4] * p.scale, acc_q[a][e * 4 + 1] * p.scale, acc_q[a][e * 4 + 2] * p.scale,
                     acc_q[a][e * 4 + 3] * p.scale);
        }
      }
    }
  }
  float* dkb = dk + static_cast<long long>(bh) * p.kv_bh_stride;
  float* dvb = dv + static_cast<long long>(bh) * p.kv_bh_stride;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = col0 + ty * 4 + a;
    if (r >= p.n_kv) continue;
#pragma unroll
    for (int e = 0; e < kCols; ++e) {
      const int c = tx * 4 + 64 * e;
      if (c < p.d) {
        *reinterpret_cast<float4*>(dkb + static_cast<long long>(r) * p.d + c) =
            make_float4(acc_k[a][e * 4] * p.scale, acc_k[a][e * 4 + 1] * p.scale, acc_k[a][e * 4 + 2] * p.scale,
                        acc_k[a][e * 4 + 3] * p.scale);
        *reinterpret_cast<float4*>(dvb + static_cast<long long>(r) * p.d + c) =
            make_float4(acc_v[a][e * 4], acc_v[a][e * 4 + 1], acc_v[a][e * 4 + 2], acc_v[a][e * 4 + 3]);
      }
    }
  }
}

static int check_shape_f32(const fa_sm100_shape* s, F32Params* p) {
  if (!s) return FA_SM100_EINVAL_PTR;
  if (s->dtype != FA_SM100_DTYPE_F32) return FA_SM100_EINVAL_DTYPE;
  if (s->d < 4 || s->d > 128 || (s->d % 4)) return FA_SM100_EINVAL_HEADDIM;
  if (s->bh <= 0 || s->n_q <= 0 || s->n_kv <= 0) return FA_SM100_EINVAL_SHAPE;
  if (s->n_q > (1ll << 30) || s->n_kv > (1ll << 30) || s->bh > 65535) return FA_SM100_EINVAL_SHAPE;
  if (!(s->softmax_scale > 0.f) || !std::isfinite(s->softmax_scale)) return FA_SM100_EINVAL_SCALE;
  const long long diag = s->q_row0 - s->kv_col0;
  if (diag > (1ll << 30) || diag < -(1ll << 30)) return FA_SM100_EINVAL_SHAPE;
  p->n_q = static_cast<int>(s->n_q);
  p->n_kv = static_cast<int>(s->n_kv);
  p->d = s->d;
  p->causal = s->causal ? 1 : 0;
  p->diag = static_cast<int>(diag);
  p->scale = s->softmax_scale;
  p->q_bh_stride = s->q_bh_stride ? s->q_bh_stride : s->n_q * s->d;
  p->kv_bh_stride = s->kv_bh_stride ? s->kv_bh_stride : s->n_kv * s->d;
  p->lse_bh_stride = s->lse_bh_stride ? s->lse_bh_stride : s->n_q;
  if (p->q_bh_stride < s->n_q * s->d || p->kv_bh_stride < s->n_kv * s->d || p->lse_bh_stride < s->n_q)
    return FA_SM100_EINVAL_SHAPE;
  if ((p->q_bh_stride % 4) || (p->kv_bh_stride % 4)) return FA_SM100_EINVAL_SHAPE;
  return FA_SM100_OK;
}

template <typename K>
static int set_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? FA_SM100_OK
                                                                                                         : FA_SM100_ELAUNCH;
}

}  // namespace fa

extern "C" int fa_sm100_fwd_f32(const fa_sm100_shape* s, const float* q, const float* k, const float* v, float* o,
                                float* lse, void* stream) {
  fa::F32Params p;
  int rc = fa::check_shape_f32(s, &p);
  if (rc) return rc;
  if (!fa::aligned16(q) || !fa::aligned16(k) || !fa::aligned16(v) || !fa::aligned16(o) || lse == nullptr)
    return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>((p.n_q + fa::kF32Tile - 1) / fa::kF32Tile), static_cast<unsigned>(s->bh));
  if (p.d <= 64) {
    constexpr int kSmem = (3 * 64 * (64 + 4) + 64 * 65) * 4;
    if ((rc = fa::set_smem(fa::fa_f32_fwd_kernel<64>, kSmem))) return rc;
    fa::fa_f32_fwd_kernel<64><<<grid, 128, kSmem, st>>>(q, k, v, o, lse, p);
  } else {
    constexpr int kSmem = (3 * 64 * (128 + 4) + 64 * 65) * 4;
    if ((rc = fa::set_smem(fa::fa_f32_fwd_kernel<128>, kSmem))) return rc;
    fa::fa_f32_fwd_kernel<128><<<grid, 128, kSmem, st>>>(q, k, v, o, lse, p);
  }
  return fa::launch_status();
}

extern "C" int fa_sm100_bwd_f32(const fa_sm100_shape* s, const float* q, const float* k, const float* v, const float* o,
                                const float* d_o, const float* lse, float* delta_ws, float* dq, float* dk, float* dv,
                                void* stream) {
  fa::F32Params p;
  int rc = fa::check_shape_f32(s, &p);
  if (rc) return rc;
  if (!fa::aligned16(q) || !fa::aligned16(k) || !fa::aligned16(v) || !fa::aligned16(o) || !fa::aligned16(d_o) ||
      lse == nullptr || delta_ws == nullptr || !fa::aligned16(dq) || !fa::aligned16(dk) || !fa::aligned16(dv))
    return FA_SM100_EINVAL_PTR;
  if ((rc = fa::check_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = s->bh * s->n_q;
  fa::fa_f32_delta_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, st>>>(o, d_o, delta_ws, s->bh, s->n_q,
                                                                                           p.d, p.q_bh_stride);
  fa::fa_f32_zero_kernel<<<148 * 8, 256, 0, st>>>(dq, s->bh, s->n_q * static_cast<long long>(p.d), p.q_bh_stride);
  const dim3 grid(static_cast<unsigned>((p.n_kv + fa::kF32Tile - 1) / fa::kF32Tile), static_cast<unsigned>(s->bh));
  if (p.d <= 64) {
    constexpr int kSmem = (4 * 64 * (64 + 4) + 2 * 64 * 65 + 128) * 4;
    if ((rc = fa::set_smem(fa::fa_f32_bwd_kernel<64>, kSmem))) return rc;
    fa::fa_f32_bwd_kernel<64><<<grid, 256, kSmem, st>>>(q, k, v, d_o, lse, delta_ws, dq, dk, dv, p);
  } else {
    constexpr int kSmem = (4 * 64 * (128 + 4) + 2 * 64 * 65 + 128) * 4;
    if ((rc = fa::set_smem(fa::fa_f32_bwd_kernel<128>, kSmem))) return rc;
    fa::fa_f32_bwd_kernel<128><<<grid, 256, kSmem, st>>>(q, k, v, d_o, lse, delta_ws, dq, dk, dv, p);
  }
  return fa::launch_status();
}
