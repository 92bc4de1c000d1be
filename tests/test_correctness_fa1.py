"""GPU parity tests for the FA1 entry point — adapted from the reference's tests/test_correctness_fa1.py
(same shapes, seeds and tolerances) with the corrected causal oracle (oracle/attention_oracle.py; the reference's own
``reference_attention(causal=True)`` masks the wrong axes).  Everything goes through the public wrappers
-> flashattention_lab_cuda shim -> C ABI -> sm_100a kernels; the oracle runs on the CPU copy of the same inputs."""
import pytest
import torch

from fa1.cuda.impl import fa1_cuda
from fa1.op import fa1_attention
from fa1.spec import pick_fa1_spec
from oracle.attention_oracle import dense_backward, dense_forward
from tests.utils import LSE_TOL, assert_allclose, dtype_tolerances, flatten_lse, flatten_output, make_qkv

pytestmark = pytest.mark.gpu
EXTRA = ()  # trailing positional args of fa1_cuda after `spec`


def _oracle_fwd(q, k, v, causal, scale):
    return dense_forward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), causal=causal, softmax_scale=scale)


# reference tests/test_correctness_fa1.py "torch forward" shapes, run through the CUDA path instead
@pytest.mark.parametrize("shape", [(1, 2, 16, 32), (2, 1, 33, 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fa1_cuda_forward_matches_reference(shape, causal, merge_heads, dtype, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(0)
    batch, heads, seqlen, head_dim = shape
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    scale = head_dim ** -0.5
    o, lse = fa1_cuda(q, k, v, causal, scale, pick_fa1_spec(head_dim), *EXTRA)
    assert o.shape == q.shape and o.dtype == dtype and lse.dtype == torch.float32
    assert lse.shape == q.shape[:-1]
    o_ref, lse_ref = _oracle_fwd(q, k, v, causal, scale)
    assert_allclose(flatten_output(o), flatten_output(o_ref), **dtype_tolerances(dtype))
    assert_allclose(flatten_lse(lse), flatten_lse(lse_ref), **LSE_TOL)


# reference tests/test_correctness_fa1.py::test_fa1_cuda_backward_matches_reference
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [True, False])
def test_fa1_cuda_backward_matches_reference(causal, merge_heads, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(3)
    dtype = torch.float16
    batch, heads, seqlen, head_dim = 1, 2, 24, 64
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, _ = fa1_cuda(q, k, v, causal, scale, pick_fa1_spec(head_dim), *EXTRA)
    do = torch.randn_like(o)
    (o * do).sum().backward()
    dq_ref, dk_ref, dv_ref, _, _ = dense_backward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(),
                                                  causal, scale)
    tol = dtype_tolerances(dtype)
    assert_allclose(q.grad, dq_ref, **tol)
    assert_allclose(k.grad, dk_ref, **tol)
    assert_allclose(v.grad, dv_ref, **tol)


# the extra fp32 "torch backward" shape of the reference suite (head dim 32), as 16-bit through the CUDA path
@pytest.mark.parametrize("causal", [False, True])
def test_fa1_cuda_backward_odd_head_dim(causal, cuda_extension_available):
    torch.manual_seed(1)
    batch, heads, seqlen, head_dim = 1, 2, 12, 32
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.bfloat16, merge_heads=True)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, lse = fa1_cuda(q, k, v, causal, scale, pick_fa1_spec(head_dim), *EXTRA)
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


# reference test_fa1_backend_consistency: shape/seed kept; the only other "backend" left is the oracle
def test_fa1_attention_entry_point_consistency(cuda_extension_available):
    torch.manual_seed(4)
    batch, heads, seqlen, head_dim = 1, 2, 24, 32
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.float16, merge_heads=False)
    o_auto, lse_auto = fa1_attention(q, k, v, causal=True)  # default scale d**-0.5, backend="auto"
    o_cuda, lse_cuda = fa1_attention(q, k, v, causal=True, softmax_scale=head_dim ** -0.5, backend="cuda")
    assert o_auto.shape == (batch, heads, seqlen, head_dim) and lse_auto.shape == (batch, heads, seqlen)
    assert torch.equal(o_auto, o_cuda) and torch.equal(lse_auto, lse_cuda)
    o_ref, lse_ref = _oracle_fwd(q, k, v, True, head_dim ** -0.5)
    assert_allclose(flatten_output(o_auto), flatten_output(o_ref), **dtype_tolerances(torch.float16))
    assert_allclose(flatten_lse(lse_auto), flatten_lse(lse_ref), **LSE_TOL)
