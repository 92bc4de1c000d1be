/*
 * fa_sm100_probes.h -- C ABI of libfa_sm100_probes.so, the DEBUG library next to libfa_sm100.so: descriptor bring-up
 * self-tests and hardware rate probes used by tools/ and one GPU test.  Not part of the product ABI (include/fa_sm100.h)
 * and not part of the reference's surface.  Error codes are those of fa_sm100.h.
 */
#ifndef FA_SM100_PROBES_H_
#define FA_SM100_PROBES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Bring-up self-test for the hand-encoded UMMA/TMA descriptors: one 128x128x128 MMA through each operand path the
 * attention kernels use.  mode 0: D = A B^T (A,B K-major smem)   1: D = A B (B MN-major smem)
 *                         2: D = A B (A from TMEM, B MN-major)   3: D = A^T B (A,B MN-major smem)
 * a, b: 128x128 row-major `dtype`; out: 128x128 fp32.  Not part of the reference surface.
 * CTA-pair forms (cluster of 2, cta_group::2, M = 256; a and out have 256 rows, b stays 128x128):
 *                         4: D = A B^T (A,B K-major smem, B rows split across the pair)
 *                         5: D = A B (A from TMEM, B MN-major, B columns split across the pair)
 * e4m3 forms (kind::f8f6f4; a, b are 128x128 BYTES of e4m3, `dtype` is ignored):
 *                         6: D = A B^T (A, B K-major smem)   7: D = A B (A from TMEM, B MN-major smem)
 */
int fa_sm100_probe_umma(int mode, int32_t dtype, const void* a, const void* b, float* out, void* stream);

/*
 * Tensor-core issue-rate probe (measurement aid, not part of the reference surface): `ctas` CTAs (CTA pairs when
 * `pair`) each stream `groups` bf16 products of shape (128 per CTA) x n x 128 -- eight UMMAs per product -- from
 * fixed operands: A from shared memory or, with `a_from_tmem`, from TMEM with B MN-major (n <= 128 then).
 * The caller times the launch: FLOPs = ctas * groups * 2 * 128 * n * 128.
 */
int fa_sm100_probe_mma_rate(int pair, int a_from_tmem, int n, int groups, int ctas, void* stream);

/*
 * L2 reduce-add rate probe (measurement aid): the backward's dQ accumulation traffic with no compute around it.
 * slices * nkt CTAs; CTA (slice, j) reduce-adds a 128x128 fp32 tile of ones into each of the nqt row tiles of
 * acc[slice] (acc: slices x (nqt*128) x 128 fp32).  flags: 1 = start at tile j (rotated walk), 2 = red.global.v4 from
 * registers instead of TMA reduce from shared memory, 4 = one CTA per SM.  Afterwards every element of acc has grown
 * by nkt.  Bytes reduced = slices * nkt * nqt * 65536.
 */
int fa_sm100_probe_reduce_rate(float* acc, int slices, int nqt, int nkt, int flags, void* stream);

/*
 * MUFU exp2 rate probe: `ctas` CTAs of 256 threads each run `iters` rounds of 8 independent exp2 chains.
 * mode 0: ex2.approx.ftz.f32   1: ex2.approx.ftz.f16x2   2: ex2.approx.ftz.bf16x2 (two results per instruction).
 * The caller times the launch; `sink` is a one-float device buffer that is never written.
 */
int fa_sm100_probe_ex2_rate(int mode, int iters, int ctas, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FA_SM100_PROBES_H_ */
