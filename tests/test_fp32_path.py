"""GPU: fp32 inputs (BASELINE config C1 and the reference suite's fp32 cases, tests/test_correctness_fa{1,2,3}.py with
tests/utils.py:31-36 tolerances rtol = atol = 1e-4, LSE 1e-3) through the public entry points -> C ABI
(fa_sm100_fwd_f32 / bwd_f32) against the fp32 oracle."""
import pytest
import torch

import flashattention_lab_cuda as ext
from fa1 import fa1_attention
from fa2 import fa2_attention
from fa3 import fa3_attention
from oracle.attention_oracle import dense_backward_fp32, error_report

pytestmark = pytest.mark.gpu
TOL = 1e-4  # reference tests/utils.py:36


def _check(got, want, tol, name):
    rep = error_report(got, want, tol, tol)
    assert rep["violations"] == 0, f"{name}: {rep}"


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("api", [fa1_attention, fa2_attention, fa3_attention])
def test_baseline_c1_fp32(api, causal):
    """B2 H4 N512 d64 fp32, forward + backward through the public API with autograd."""
    torch.manual_seed(0)
    b, h, n, d = 2, 4, 512, 64
    q, k, v = (torch.randn(b, h, n, d, device="cuda", dtype=torch.float32).requires_grad_(True) for _ in range(3))
    do = torch.randn(b, h, n, d, device="cuda")
    o, lse = api(q, k, v, causal=causal, backend="cuda")
    assert o.dtype == torch.float32 and lse.dtype == torch.float32 and o.shape == q.shape and lse.shape == (b, h, n)
    o.backward(do)
    dq_r, dk_r, dv_r, o_r, lse_r = dense_backward_fp32(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(), causal)
    _check(o, o_r, TOL, "o")
    _check(lse, lse_r, 1e-3, "lse")
    for name, got, want in (("dq", q.grad, dq_r), ("dk", k.grad, dk_r), ("dv", v.grad, dv_r)):
        _check(got, want, TOL, name)
        assert got.dtype == torch.float32


@pytest.mark.parametrize("bh,n_q,n_kv,d,causal", [(3, 33, 33, 32, True), (2, 24, 24, 48, False), (1, 200, 457, 128, False),
                                                  (2, 130, 130, 40, True), (2, 1000, 1000, 64, True), (1, 65, 64, 100, True),
                                                  (2, 77, 300, 22, False)])
def test_fp32_shapes(bh, n_q, n_kv, d, causal):
    torch.manual_seed(n_q + d)
    q, do = (torch.randn(bh, n_q, d, device="cuda") for _ in range(2))
    k, v = (torch.randn(bh, n_kv, d, device="cuda") for _ in range(2))
    scale = d ** -0.5
    q_row0 = n_kv - n_q if causal and n_kv > n_q else 0
    if d % 4 == 0:
        o, lse = ext.fwd_f32_raw(q, k, v, causal, scale, q_row0=q_row0)
        dq, dk, dv = ext.bwd_f32_raw(q, k, v, o, do, lse, causal, scale, q_row0=q_row0)
    else:  # the public functions pad the head dim to a multiple of 4 (square shapes only: the reference's contract)
        k, v = k[:, :n_q].contiguous(), v[:, :n_q].contiguous()
        n_kv, q_row0 = n_q, 0
        o, lse = ext.forward(q, k, v, causal, scale, 128, 128)
        dq, dk, dv = ext.backward(q, k, v, o, do, lse, causal, scale, 128, 128)
    dq_r, dk_r, dv_r, o_r, lse_r = dense_backward_fp32(q.cpu(), k.cpu(), v.cpu(), do.cpu(), causal, scale, q_row0, 0)
    for name, got, want, tol in (("o", o, o_r, TOL), ("lse", lse, lse_r, 1e-3), ("dq", dq, dq_r, 2e-4), ("dk", dk, dk_r, 2e-4),
                                 ("dv", dv, dv_r, 2e-4)):
        assert got.shape == want.shape
        _check(got, want, tol, name)


def test_fp32_rows_without_visible_keys_and_errors():
    torch.manual_seed(2)
    q, k, v = (torch.randn(2, 96, 64, device="cuda") for _ in range(3))
    o, lse = ext.fwd_f32_raw(q, k, v, True, 0.125, kv_col0=40)  # rows 0..39 see nothing
    assert torch.all(o[:, :40] == 0) and torch.all(torch.isinf(lse[:, :40]) & (lse[:, :40] < 0))
    assert torch.isfinite(o).all() and torch.isfinite(lse[:, 40:]).all()
    with pytest.raises(NotImplementedError):
        big = torch.randn(1, 16, 132, device="cuda")
        ext.forward(big, big, big, False, 0.1, 128, 128)
    with pytest.raises(NotImplementedError):
        ext.fwd_raw(q, k, v, False, 0.125)  # the tcgen05 entry point itself stays 16-bit only
