"""Test helpers with the reference's conventions (reference tests/utils.py:7-36)."""
from __future__ import annotations

import torch

from common.utils import merge_bh


def make_qkv(batch, heads, seqlen, head_dim, device, dtype, merge_heads=False):
    shape = (batch, heads, seqlen, head_dim)
    q = torch.randn(shape, device=device, dtype=dtype)
    k = torch.randn(shape, device=device, dtype=dtype)
    v = torch.randn(shape, device=device, dtype=dtype)
    if merge_heads:
        q, k, v = (merge_bh(t)[0] for t in (q, k, v))
    return q, k, v


def flatten_output(x):
    return x.reshape(-1, x.shape[-2], x.shape[-1]) if x.dim() == 4 else x


def flatten_lse(x):
    return x.reshape(-1, x.shape[-1]) if x.dim() == 3 else x


def dtype_tolerances(dtype):
    # reference tests/utils.py:31-36
    if dtype in (torch.float16, torch.bfloat16):
        return {"rtol": 5e-2, "atol": 5e-2}
    return {"rtol": 1e-4, "atol": 1e-4}


LSE_TOL = {"rtol": 1e-3, "atol": 1e-3}  # reference tests/test_correctness_fa2.py:33


def assert_allclose(actual, expected, rtol, atol, msg=None):
    torch.testing.assert_close(actual.detach().float().cpu(), expected.detach().float().cpu(), rtol=rtol, atol=atol,
                               msg=msg)
