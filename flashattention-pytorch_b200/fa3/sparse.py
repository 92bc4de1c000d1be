"""Block-sparse attention with dropout: the operator behind the reference's stand-alone module
(``src/fa3/torch/flashattention_pytorch.py``) on the sm_100a kernels.

Semantics are those of that module's dense branch (``:80-87``: ``scores.masked_fill(mask == 0, -inf)`` ->
``softmax`` -> ``dropout`` -> ``@ v``) with the tile skip of its block-sparse branch (``:124``:
``block_sparse_mask[i, j] == 0`` tiles are not computed).  Block size is 128, the module's default (``:19``).
The module's own tiled branch applies dropout to the UN-normalised exponentials and folds the dropped values into its
running denominator (``:146-160``), which is not an unbiased dropout of the attention probabilities and whose backward
(``:381-430``) re-normalises per tile; that branch is deliberately not mirrored (deviations).

    o, lse = fa3_block_sparse_attention(q, k, v, block_sparse_mask=m, dropout_p=0.1, causal=True)

``q, k, v``: (B, H, N, d) or (BH, N, d) CUDA fp16/bf16.  ``block_sparse_mask``: (ceil(N/128), ceil(N/128)) shared by
every batch*head, or one per batch*head with leading dims (B, H) / (BH,); nonzero = compute.  Dropout bits come from
Philox keyed by ``seed`` (drawn from torch's default generator when omitted) and the element coordinates; the backward
regenerates them.  Differentiable; ``lse`` does not depend on dropout.  No fallback: raises without the CUDA library."""
from __future__ import annotations

import torch

from common.autograd_cuda import load_ext, require_cuda
from common.utils import merge_bh, split_bh, split_bh_lse


class _BlockSparseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, causal, softmax_scale, dropout_p, seed):
        ext = load_ext()
        require_cuda(q, k, v)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = ext.fwd_ex_raw(q, k, v, bool(causal), float(softmax_scale), block_mask=mask, dropout_p=dropout_p,
                                seed=seed)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.mask = mask
        ctx.meta = (bool(causal), float(softmax_scale), float(dropout_p), int(seed))
        return o, lse

    @staticmethod
    def backward(ctx, do, dlse):  # dlse is ignored, like the reference's CUDA wrappers
        ext = load_ext()
        q, k, v, o, lse = ctx.saved_tensors
        causal, scale, p, seed = ctx.meta
        dq, dk, dv = ext.bwd_ex_raw(q, k, v, o, do.contiguous().to(q.dtype), lse, causal, scale, block_mask=ctx.mask,
                                    dropout_p=p, seed=seed)
        return dq, dk, dv, None, None, None, None, None


def fa3_block_sparse_attention(q, k, v, block_sparse_mask=None, dropout_p=0.0, causal=False, softmax_scale=None,
                               training=True, seed=None):
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    p = float(dropout_p) if training else 0.0
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
    qb, bh_shape = merge_bh(q)
    kb, _ = merge_bh(k)
    vb, _ = merge_bh(v)
    mask = block_sparse_mask
    if mask is not None and mask.dim() == 4:  # (B, H, q_blocks, k_blocks) -> (BH, q_blocks, k_blocks)
        mask = mask.reshape(-1, mask.shape[-2], mask.shape[-1])
    if mask is not None:
        mask = mask.to(qb.device)
    o, lse = _BlockSparseFn.apply(qb, kb, vb, mask, causal, softmax_scale, p, seed)
    return split_bh(o, bh_shape), split_bh_lse(lse, bh_shape)
