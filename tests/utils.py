"""Test helpers that follow the reference's conventions (reference tests/utils.py:7-36): seeded standard-normal
q/k/v of shape (B, H, N, d) optionally merged to (B*H, N, d), outputs compared after flattening the head axes, and the
per-dtype tolerances the reference suite uses."""
from __future__ import annotations

import torch

from common.utils import merge_bh

# reference tests/utils.py:31-36 -- 16-bit inputs are compared at 5e-2, everything else at 1e-4
_TOLERANCE = {torch.float16: 5e-2, torch.bfloat16: 5e-2}
LSE_TOL = {"rtol": 1e-3, "atol": 1e-3}  # reference tests/test_correctness_fa2.py:33


def make_qkv(batch, heads, seqlen, head_dim, device, dtype, merge_heads=False):
    """Three independent randn draws, in q, k, v order (the order fixes the values for a given seed)."""
    drawn = [torch.randn((batch, heads, seqlen, head_dim), device=device, dtype=dtype) for _ in "qkv"]
    return tuple(merge_bh(t)[0] for t in drawn) if merge_heads else tuple(drawn)


def _collapse_leading(x, keep):
    return x if x.dim() == keep else x.reshape(-1, *x.shape[-(keep - 1):])


def flatten_output(x):
    return _collapse_leading(x, 3)  # (B, H, N, d) -> (B*H, N, d)


def flatten_lse(x):
    return _collapse_leading(x, 2)  # (B, H, N) -> (B*H, N)


def dtype_tolerances(dtype):
    tol = _TOLERANCE.get(dtype, 1e-4)
    return {"rtol": tol, "atol": tol}


def assert_allclose(actual, expected, rtol, atol, msg=None):
    torch.testing.assert_close(actual.detach().float().cpu(), expected.detach().float().cpu(), rtol=rtol, atol=atol,
                               msg=msg)
