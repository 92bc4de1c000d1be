"""CPU oracle for the attention forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this module; the product path (``flashattention-pytorch_b200/``) never does and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
  * outputs of the reference's own Python implementation (``fa1_forward_torch`` / ``fa1_backward_torch`` /
    ``fa3_forward_torch`` / non-causal ``reference_attention`` + autograd) generated in the build container by
    ``oracle/make_golden.py`` and committed under ``tests/golden/``;
  * and, when present, the reference's csrc compiled for CPU (``oracle/_ref``, built by ``oracle/build_ref.py``).
The reference holds no golden vectors of its own: its tests compute expectations on the fly.
Round-2 additions and their pins:
  * ``expand_block_mask`` / ``dense_ext_backward_fp32`` WITHOUT dropout: PINNED against outputs of the reference's
    stand-alone block-sparse module (``tests/golden/ref_block_sparse.npz``, block sizes 32 and 64).
  * dropout (``philox4x32_7`` / ``dropout_keep_mask``): PARITY UNPINNED against the reference — its dropout draws from
    torch's global generator and its tiled branch drops un-normalised exponentials, so no bit pattern of it can be
    reproduced.  The generator itself is pinned to the published Philox4x32-10 known-answer vectors.
  * FP8 (``fp8_quantize_dequantize`` / ``fp8_forward_oracle``): PARITY UNPINNED against the reference — its fp8 emulation
    is numerically broken; this oracle is fp32 attention on quantise -> dequantise inputs, defined here.
    Only its per-block scales are pinned (= the reference's ``_block_absmax_scale`` / 448, ``ref_fp8_helpers.npz``).

Two restatements of the same operator, fp32 arithmetic on whatever dtype the inputs have (inputs are up-cast, as the
reference does):

``dense_forward`` / ``dense_backward``
    follows reference ``src/common/correctness.py:5-34``: ``softmax(Q K^T * scale [+ causal]) V``, ``o.to(q.dtype)``,
    ``lse = logsumexp``; gradients by autograd through the fp32 graph, returned in the input dtype.  The causal mask
    is the rule of reference ``src/fa1/torch/impl.py:18-24`` (key ``c`` hidden from query ``r`` iff ``c > r``) applied
    to the last two dims — the reference's own ``reference_attention(causal=True)`` applies it to the wrong axes
    (SURVEY.md defect D1), so that one call site is intentionally NOT followed.

``blocked_forward`` / ``blocked_backward``
    follows reference ``src/fa1/torch/impl.py:26-68`` and ``:70-115``: the tiled online-softmax recurrence and the
    KV-outer backward with ``P = exp(S - lse)``; vectorised over the batch*head dimension so it is usable at a few
    thousand rows.  Supports the ring-attention generalisation (``q_row0`` / ``kv_col0`` offsets, n_q != n_kv).
"""
from __future__ import annotations

import math

import torch

NEG_INF = float("-inf")


def _as_bh(x):
    """(B,H,N,D) or (BH,N,D) -> (BH,N,D), plus the (B,H) to restore (reference src/common/utils.py:3-7)."""
    if x.dim() == 4:
        return x.reshape(x.shape[0] * x.shape[1], x.shape[2], x.shape[3]), (x.shape[0], x.shape[1])
    return x, None


def visible_mask(n_q, n_kv, q_row0=0, kv_col0=0, device="cpu"):
    """True where key c is visible to query r under the causal rule  kv_col0 + c <= q_row0 + r
    (reference src/fa1/torch/impl.py:18-24 with both offsets zero)."""
    r = torch.arange(n_q, device=device)[:, None] + q_row0
    c = torch.arange(n_kv, device=device)[None, :] + kv_col0
    return c <= r


# ----------------------------------------------------------------------------------------------------------------------
# dense
# ----------------------------------------------------------------------------------------------------------------------
def dense_forward(q, k, v, causal=False, softmax_scale=None, q_row0=0, kv_col0=0):
    """reference src/common/correctness.py:5-24.  Returns (o in q.dtype, lse fp32)."""
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    qb, bh_shape = _as_bh(q)
    kb, _ = _as_bh(k)
    vb, _ = _as_bh(v)
    s = torch.matmul(qb.float(), kb.float().transpose(-2, -1)) * softmax_scale
    if causal:
        s = s.masked_fill(~visible_mask(qb.shape[1], kb.shape[1], q_row0, kv_col0, s.device), NEG_INF)
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)[..., None])  # all-masked row -> 0
    o = torch.matmul(p, vb.float()).to(q.dtype)
    if bh_shape is not None:
        o = o.reshape(*bh_shape, *o.shape[1:])
        lse = lse.reshape(*bh_shape, lse.shape[-1])
    return o, lse


def dense_backward(q, k, v, do, causal=False, softmax_scale=None, q_row0=0, kv_col0=0):
    """reference src/common/correctness.py:26-34: autograd through the dense fp32 graph.
    Returns (dq, dk, dv, o, lse); grads in the input dtype."""
    qr = q.detach().clone().requires_grad_(True)
    kr = k.detach().clone().requires_grad_(True)
    vr = v.detach().clone().requires_grad_(True)
    o, lse = dense_forward(qr, kr, vr, causal, softmax_scale, q_row0, kv_col0)
    o.backward(do.to(o.dtype))
    return qr.grad, kr.grad, vr.grad, o.detach(), lse.detach()


def dense_backward_fp32(q, k, v, do, causal=False, softmax_scale=None, q_row0=0, kv_col0=0):
    """Closed-form fp32 gradients (no rounding of O or the grads to the input dtype): the target a 16-bit kernel
    with fp32 accumulation should be compared with.  Same maths as reference csrc/fa1/fa1_bwd.cu:57,96-104."""
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    qb, bh_shape = _as_bh(q)
    kb, _ = _as_bh(k)
    vb, _ = _as_bh(v)
    dob, _ = _as_bh(do)
    qf, kf, vf, dof = qb.float(), kb.float(), vb.float(), dob.float()
    s = torch.matmul(qf, kf.transpose(-2, -1)) * softmax_scale
    if causal:
        s = s.masked_fill(~visible_mask(qf.shape[1], kf.shape[1], q_row0, kv_col0, s.device), NEG_INF)
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)[..., None])
    o = torch.matmul(p, vf)
    delta = (dof * o).sum(-1, keepdim=True)
    dv = torch.matmul(p.transpose(-2, -1), dof)
    dp = torch.matmul(dof, vf.transpose(-2, -1))
    ds = p * (dp - delta)
    dq = torch.matmul(ds, kf) * softmax_scale
    dk = torch.matmul(ds.transpose(-2, -1), qf) * softmax_scale
    outs = [dq, dk, dv, o, lse]
    if bh_shape is not None:
        outs = [t.reshape(*bh_shape, *t.shape[1:]) for t in outs]
    return tuple(outs)


# ----------------------------------------------------------------------------------------------------------------------
# blocked (tiled) — the algorithm the kernels implement
# ----------------------------------------------------------------------------------------------------------------------
def blocked_forward(q, k, v, causal, softmax_scale, br=128, bc=128, q_row0=0, kv_col0=0):
    """reference src/fa1/torch/impl.py:26-68 on (BH,N,D) tensors, all slices at once.
    Tile skip: a KV tile is skipped iff its first key is hidden from the tile's last query (``:15-16,41-42``);
    the element mask is applied wherever a tile straddles the diagonal."""
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    o = torch.empty((bh, n_q, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((bh, n_q), dtype=torch.float32, device=q.device)
    for r0 in range(0, n_q, br):
        r1 = min(r0 + br, n_q)
        qi = q[:, r0:r1].float()
        m = torch.full((bh, r1 - r0), NEG_INF)
        l = torch.zeros((bh, r1 - r0))
        acc = torch.zeros((bh, r1 - r0, d))
        for c0 in range(0, n_kv, bc):
            if causal and kv_col0 + c0 > q_row0 + r1 - 1:
                break
            c1 = min(c0 + bc, n_kv)
            s = torch.matmul(qi, k[:, c0:c1].float().transpose(-2, -1)) * softmax_scale
            if causal and kv_col0 + c1 - 1 > q_row0 + r0:
                rr = torch.arange(r0, r1)[:, None] + q_row0
                cc = torch.arange(c0, c1)[None, :] + kv_col0
                s = s.masked_fill(cc > rr, NEG_INF)
            m_new = torch.maximum(m, s.amax(dim=-1))
            m_safe = torch.where(torch.isinf(m_new), torch.zeros_like(m_new), m_new)
            p = torch.exp(s - m_safe[..., None])
            alpha = torch.exp(m - m_safe)
            l = alpha * l + p.sum(-1)
            acc = alpha[..., None] * acc + torch.matmul(p, v[:, c0:c1].float())
            m = m_new
        safe_l = torch.where(l > 0, l, torch.ones_like(l))
        o[:, r0:r1] = (acc / safe_l[..., None]).to(q.dtype)
        lse[:, r0:r1] = torch.where(l > 0, m + torch.log(safe_l), torch.full_like(l, NEG_INF))
    return o, lse


def blocked_backward(q, k, v, o, do, lse, causal, softmax_scale, br=128, bc=128, q_row0=0, kv_col0=0,
                     out_dtype=None):
    """reference src/fa1/torch/impl.py:70-115: KV-outer / Q-inner, P recomputed from the saved lse.
    Returns (dq, dk, dv) in ``out_dtype`` (default: input dtype, like the reference's ``.to(q.dtype)``)."""
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    dq = torch.zeros((bh, n_q, d))
    dk = torch.zeros((bh, n_kv, d))
    dv = torch.zeros((bh, n_kv, d))
    delta = (do.float() * o.float()).sum(-1)
    lse_safe = torch.where(torch.isinf(lse), torch.zeros_like(lse), lse).float()
    for c0 in range(0, n_kv, bc):
        c1 = min(c0 + bc, n_kv)
        kj, vj = k[:, c0:c1].float(), v[:, c0:c1].float()
        for r0 in range(0, n_q, br):
            r1 = min(r0 + br, n_q)
            if causal and kv_col0 + c0 > q_row0 + r1 - 1:
                continue
            qi, doi = q[:, r0:r1].float(), do[:, r0:r1].float()
            s = torch.matmul(qi, kj.transpose(-2, -1)) * softmax_scale
            if causal and kv_col0 + c1 - 1 > q_row0 + r0:
                rr = torch.arange(r0, r1)[:, None] + q_row0
                cc = torch.arange(c0, c1)[None, :] + kv_col0
                s = s.masked_fill(cc > rr, NEG_INF)
            p = torch.exp(s - lse_safe[:, r0:r1, None])
            dv[:, c0:c1] += torch.matmul(p.transpose(-2, -1), doi)
            dp = torch.matmul(doi, vj.transpose(-2, -1))
            ds = p * (dp - delta[:, r0:r1, None])
            dq[:, r0:r1] += torch.matmul(ds, kj) * softmax_scale
            dk[:, c0:c1] += torch.matmul(ds.transpose(-2, -1), qi) * softmax_scale
    out_dtype = out_dtype or q.dtype
    return dq.to(out_dtype), dk.to(out_dtype), dv.to(out_dtype)


def merge_partials(o_a, lse_a, o_b, lse_b):
    """Log-sum-exp merge of two attention partials over disjoint key sets (the ring-attention combine; derived from
    the online-softmax update of reference src/fa1/torch/impl.py:53-62)."""
    lse = torch.logaddexp(lse_a, lse_b)
    safe = torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)
    wa = torch.exp(lse_a - safe)[..., None]
    wb = torch.exp(lse_b - safe)[..., None]
    return (wa * o_a.float() + wb * o_b.float()), lse


# ----------------------------------------------------------------------------------------------------------------------
# error report (SURVEY.md §8d "Parity report")
# ----------------------------------------------------------------------------------------------------------------------
def error_report(actual, expected, rtol, atol):
    a, e = actual.detach().float().cpu(), expected.detach().float().cpu()
    finite = torch.isfinite(e)
    same_inf = (~finite) & (a == e)
    diff = torch.where(finite, (a - e).abs(), torch.where(same_inf, torch.zeros_like(a), torch.full_like(a, math.inf)))
    big = finite & (e.abs() > 1e-2)
    rel = torch.where(big, diff / e.abs().clamp_min(1e-30), torch.zeros_like(diff))
    viol = diff > (atol + rtol * torch.where(finite, e.abs(), torch.zeros_like(e)))
    return {
        "max_abs": float(diff.max()) if diff.numel() else 0.0,
        "max_rel": float(rel.max()) if rel.numel() else 0.0,
        "violations": int(viol.sum()),
        "numel": int(diff.numel()),
    }


def relative_error(actual, expected, tile_rows=128):
    """Scale-aware companion of ``error_report`` for long sequences, where an absolute 5e-2 is about the RMS of the
    gradients themselves: ||a - e|| / ||e|| over the whole tensor and the worst such ratio over ``tile_rows``-row
    tiles of the second-to-last dimension (a late-row or single-tile accumulation error cannot hide in the total)."""
    a, e = actual.detach().float(), expected.detach().float().to(actual.device)
    total = float((a - e).norm() / e.norm().clamp_min(1e-30))
    n = a.shape[-2]
    pad = (-n) % tile_rows
    if pad:
        a = torch.nn.functional.pad(a, (0, 0, 0, pad))
        e = torch.nn.functional.pad(e, (0, 0, 0, pad))
    at = a.reshape(*a.shape[:-2], -1, tile_rows, a.shape[-1])
    et = e.reshape(*e.shape[:-2], -1, tile_rows, e.shape[-1])
    num = (at - et).flatten(-2).norm(dim=-1)
    den = et.flatten(-2).norm(dim=-1)
    ok = den > 1e-6 * float(e.norm()) / max(1, den.numel()) ** 0.5  # skip tiles that are (numerically) all zero
    worst = float((num[ok] / den[ok]).max()) if ok.any() else 0.0
    return {"rel_fro": total, "worst_tile_rel_fro": worst}


# ----------------------------------------------------------------------------------------------------------------------
# block-sparse mask + dropout
# ----------------------------------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_U32 = 0xFFFFFFFF


def philox4x32_7(c0, c1, c2, c3, k0, k1, rounds=7):
    """Philox4x32 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) on int64 tensors holding 32-bit
    values; 7 rounds is what the kernels use for dropout (csrc/ptx.cuh philox4x32_7), ``rounds=10`` reproduces the
    published known-answer vectors (tests/test_oracle_golden.py).  Returns the four output words."""
    def mulhilo(a, b):  # 32x32 -> (hi, lo) without overflowing int64: split b into 16-bit halves
        lo_part = a * (b & 0xFFFF)
        hi_part = a * (b >> 16)
        full_lo = (lo_part + ((hi_part & 0xFFFF) << 16))
        lo = full_lo & _U32
        hi = ((hi_part >> 16) + (full_lo >> 32)) & _U32
        return hi, lo

    c0, c1, c2, c3 = (x.clone() & _U32 for x in (c0, c1, c2, c3))
    for _ in range(rounds):
        hi0, lo0 = mulhilo(c0, _PHILOX_M0)
        hi1, lo1 = mulhilo(c2, _PHILOX_M1)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _U32, lo1, (hi0 ^ c3 ^ k1) & _U32, lo0
        k0, k1 = (k0 + _PHILOX_W0) & _U32, (k1 + _PHILOX_W1) & _U32
    return c0, c1, c2, c3


def dropout_keep_mask(bh, n_q, n_kv, dropout_p, seed, offset=0, q_row0=0, kv_col0=0):
    """(bh, n_q, n_kv) bool keep mask and the rescale factor, bit-for-bit what the kernels generate: one Philox call
    per 4 x 4 block of (query, key) elements, counter (query >> 2, key >> 2, slice, offset), key (seed lo, seed hi);
    the element's byte is word (query & 3), byte (key & 3); it is dropped iff byte < floor(p * 256)."""
    thr = int(float(dropout_p) * 256.0)
    if thr == 0:
        return torch.ones((bh, n_q, n_kv), dtype=torch.bool), 1.0
    qg = (torch.arange(n_q, dtype=torch.int64) + q_row0)[None, :, None]
    kg = (torch.arange(n_kv, dtype=torch.int64) + kv_col0)[None, None, :]
    sl = torch.arange(bh, dtype=torch.int64)[:, None, None]
    shape = (bh, n_q, n_kv)
    words = philox4x32_7((qg >> 2).expand(shape), (kg >> 2).expand(shape), sl.expand(shape),
                         torch.full(shape, int(offset) & _U32, dtype=torch.int64), int(seed) & _U32,
                         (int(seed) >> 32) & _U32)
    w = torch.stack(words, dim=-1).gather(-1, (qg & 3).expand(shape)[..., None]).squeeze(-1)
    byte = (w >> ((kg & 3) * 8).expand(shape)) & 0xFF
    return byte >= thr, 256.0 / (256.0 - thr)


def expand_block_mask(block_mask, n_q, n_kv, block=128):
    """(…, ceil(n_q/block), ceil(n_kv/block)) tile mask -> (…, n_q, n_kv) element mask (nonzero = visible)."""
    m = (block_mask != 0)
    m = m.repeat_interleave(block, dim=-2).repeat_interleave(block, dim=-1)
    return m[..., :n_q, :n_kv]


def dense_ext_backward_fp32(q, k, v, do, causal=False, softmax_scale=None, block_mask=None, dropout_p=0.0, seed=0,
                            offset=0, q_row0=0, kv_col0=0, block=128):
    """fp32 forward + closed-form gradients of
        O = dropout(softmax(mask(Q K^T * scale))) V
    as the dense branch of the reference's stand-alone module computes it (src/fa3/torch/flashattention_pytorch.py:80-87:
    masked_fill(mask == 0, -inf) -> softmax -> dropout -> @ v), with the block-sparse tile mask (:124) expanded to
    elements and the kernels' Philox keep mask.  Returns (dq, dk, dv, o, lse); lse ignores dropout."""
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    qf, kf, vf, dof = (t.float().cpu() for t in (q, k, v, do))
    bh, n_q, _ = qf.shape
    n_kv = kf.shape[1]
    s = torch.matmul(qf, kf.transpose(-2, -1)) * softmax_scale
    vis = torch.ones((n_q, n_kv), dtype=torch.bool)
    if causal:
        vis = visible_mask(n_q, n_kv, q_row0, kv_col0)
    vis = vis[None].expand(bh, n_q, n_kv)
    if block_mask is not None:
        bm = expand_block_mask(block_mask.cpu(), n_q, n_kv, block)
        vis = vis & (bm if bm.dim() == 3 else bm[None])
    s = s.masked_fill(~vis, NEG_INF)
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)[..., None])
    keep, rescale = dropout_keep_mask(bh, n_q, n_kv, dropout_p, seed, offset, q_row0, kv_col0)
    z = keep.float() * rescale
    pd = p * z
    o = torch.matmul(pd, vf)
    delta = (dof * o).sum(-1, keepdim=True)
    dv = torch.matmul(pd.transpose(-2, -1), dof)
    dp = torch.matmul(dof, vf.transpose(-2, -1)) * z
    ds = p * (dp - delta)
    dq = torch.matmul(ds, kf) * softmax_scale
    dk = torch.matmul(ds.transpose(-2, -1), qf) * softmax_scale
    return dq, dk, dv, o, lse


# ----------------------------------------------------------------------------------------------------------------------
# FP8 forward
# ----------------------------------------------------------------------------------------------------------------------
def fp8_quantize_dequantize(x, hadamard, seed=0, block=128):
    """What the kernels' quantisation pre-pass does to a (bh, n, 128) tensor, returned DE-quantised in fp32 (plus the
    e4m3 tensor and the per-block scales): optional incoherent processing as the reference's emulation
    (src/fa3/torch/impl.py:41-59: random sign flip, Walsh-Hadamard transform, 1/sqrt(d)) with signs from Philox bits of
    ``seed``; then per-``block``-rows absmax / 448 scales (:20-31) and round-to-nearest e4m3 (torch.float8_e4m3fn)."""
    xf = x.float().cpu()
    bh, n, d = xf.shape
    if hadamard:
        lanes = torch.arange(d // 4, dtype=torch.int64)
        zero = torch.zeros_like(lanes)
        words = philox4x32_7(lanes, zero, zero, zero, int(seed) & _U32, (int(seed) >> 32) & _U32)
        bits = torch.stack(words, dim=-1).reshape(-1) & 1  # element 4 * lane + i takes word i of lane's call
        xf = xf * (1.0 - 2.0 * bits.float())
        h = 1
        while h < d:
            y = xf.reshape(bh, n, d // (2 * h), 2, h)
            xf = torch.stack((y[..., 0, :] + y[..., 1, :], y[..., 0, :] - y[..., 1, :]), dim=-2).reshape(bh, n, d)
            h *= 2
        xf = xf * (d ** -0.5)
    nb = (n + block - 1) // block
    pad = nb * block - n
    xp = torch.nn.functional.pad(xf, (0, 0, 0, pad)).reshape(bh, nb, block * d)
    amax = xp.abs().amax(dim=-1)
    scale = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    q8 = (xp / scale[..., None]).to(torch.float8_e4m3fn)
    deq = (q8.float() * scale[..., None]).reshape(bh, nb * block, d)[:, :n]
    return deq, q8.reshape(bh, nb * block, d)[:, :n], scale


def fp8_forward_oracle(q, k, v, causal=False, softmax_scale=None, seed=0):
    """fp32 attention over the quantise -> dequantise images of Q, K (rotated) and V: the target of the e4m3 kernel up
    to its e4m3 rounding of the probabilities (not modelled here: it depends on the running row maximum)."""
    qd, _, _ = fp8_quantize_dequantize(q, True, seed)
    kd, _, _ = fp8_quantize_dequantize(k, True, seed)
    vd, _, _ = fp8_quantize_dequantize(v, False)
    return dense_forward(qd, kd, vd, causal, softmax_scale)

