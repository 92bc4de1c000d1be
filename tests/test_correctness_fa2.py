"""GPU parity tests for the FA2 entry point — adapted from the reference's tests/test_correctness_fa2.py
(same shapes, seeds and tolerances) with the corrected causal oracle (oracle/attention_oracle.py; the reference's own
``reference_attention(causal=True)`` masks the wrong axes).  Everything goes through the public wrappers
-> flashattention_lab_cuda shim -> C ABI -> sm_100a kernels; the oracle runs on the CPU copy of the same inputs."""
import pytest
import torch

from fa2.cuda.impl import fa2_cuda
from fa2.op import fa2_attention
from fa2.spec import pick_fa2_spec
from oracle.attention_oracle import dense_backward, dense_forward
from tests.utils import LSE_TOL, assert_allclose, dtype_tolerances, flatten_lse, flatten_output, make_qkv

pytestmark = pytest.mark.gpu
EXTRA = ()  # trailing positional args of fa2_cuda after `spec`


def _oracle_fwd(q, k, v, causal, scale):
    return dense_forward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), causal=causal, softmax_scale=scale)


# reference tests/test_correctness_fa2.py "torch forward" shapes, run through the CUDA path instead
@pytest.mark.parametrize("shape", [(1, 1, 24, 32), (2, 2, 33, 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fa2_cuda_forward_matches_reference(shape, causal, merge_heads, dtype, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(10)
    batch, heads, seqlen, head_dim = shape
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    scale = head_dim ** -0.5
    o, lse = fa2_cuda(q, k, v, causal, scale, pick_fa2_spec(head_dim), *EXTRA)
    assert o.shape == q.shape and o.dtype == dtype and lse.dtype == torch.float32
    assert lse.shape == q.shape[:-1]
    o_ref, lse_ref = _oracle_fwd(q, k, v, causal, scale)
    assert_allclose(flatten_output(o), flatten_output(o_ref), **dtype_tolerances(dtype))
    assert_allclose(flatten_lse(lse), flatten_lse(lse_ref), **LSE_TOL)


# reference tests/test_correctness_fa2.py::test_fa2_cuda_backward_matches_reference
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [True, False])
def test_fa2_cuda_backward_matches_reference(causal, merge_heads, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(13)
    dtype = torch.float16
    batch, heads, seqlen, head_dim = 1, 2, 32, 48
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, _ = fa2_cuda(q, k, v, causal, scale, pick_fa2_spec(head_dim), *EXTRA)
    do = torch.randn_like(o)
    (o * do).sum().backward()
    dq_ref, dk_ref, dv_ref, _, _ = dense_backward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(),
                                                  causal, scale)
    tol = dtype_tolerances(dtype)
    assert_allclose(q.grad, dq_ref, **tol)
    assert_allclose(k.grad, dk_ref, **tol)
    assert_allclose(v.grad, dv_ref, **tol)


# the extra fp32 "torch backward" shape of the reference suite (head dim 40), as 16-bit through the CUDA path
@pytest.mark.parametrize("causal", [False, True])
def test_fa2_cuda_backward_odd_head_dim(causal, cuda_extension_available):
    torch.manual_seed(11)
    batch, heads, seqlen, head_dim = 1, 2, 16, 40
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.bfloat16, merge_heads=True)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, lse = fa2_cuda(q, k, v, causal, scale, pick_fa2_spec(head_dim), *EXTRA)
    do = torch.randn_like(o)
    o.backward(do)
    dq_ref, dk_ref, dv_ref, o_ref, lse_ref = dense_backward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(),
                                                            do.cpu(), causal, scale)
    tol = dtype_tolerances(torch.bfloat16)
    assert_allclose(o, o_ref, **tol)
    assert_allclose(lse, lse_ref, **LSE_TOL)
    for got, want in ((q.grad, dq_ref), (k.grad, dk_ref), (v.grad, dv_ref)):
        assert got.shape == want.shape
        assert_allclose(got, want, **tol)


# reference test_fa2_backend_consistency: shape/seed kept; the only other "backend" left is the oracle
def test_fa2_attention_entry_point_consistency(cuda_extension_available):
    torch.manual_seed(14)
    batch, heads, seqlen, head_dim = 1, 2, 28, 32
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.float16, merge_heads=False)
    o_auto, lse_auto = fa2_attention(q, k, v, causal=True)  # default scale d**-0.5, backend="auto"
    o_cuda, lse_cuda = fa2_attention(q, k, v, causal=True, softmax_scale=head_dim ** -0.5, backend="cuda")
    assert o_auto.shape == (batch, heads, seqlen, head_dim) and lse_auto.shape == (batch, heads, seqlen)
    assert torch.equal(o_auto, o_cuda) and torch.equal(lse_auto, lse_cuda)
    o_ref, lse_ref = _oracle_fwd(q, k, v, True, head_dim ** -0.5)
    assert_allclose(flatten_output(o_auto), flatten_output(o_ref), **dtype_tolerances(torch.float16))
    assert_allclose(flatten_lse(lse_auto), flatten_lse(lse_ref), **LSE_TOL)


def test_fa2_forward_is_normalised_once(cuda_extension_available):
    """Deviation from the reference on purpose: its FA2 forward divides O by the row sum twice;
    the tests (its contract) compare against true attention, which is what this returns."""
    torch.manual_seed(10)
    q, k, v = make_qkv(1, 2, 40, 64, device="cuda", dtype=torch.float16, merge_heads=True)
    o, _ = fa2_cuda(q, k, v, False, 0.125, pick_fa2_spec(64))
    ones = torch.ones_like(v)
    o1, _ = fa2_cuda(q, k, ones, False, 0.125, pick_fa2_spec(64))
    assert_allclose(o1, torch.ones_like(o1), rtol=2e-3, atol=2e-3)  # softmax rows sum to one
