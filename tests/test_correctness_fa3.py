"""GPU parity tests for the FA3 entry point — adapted from the reference's tests/test_correctness_fa3.py
(same shapes, seeds and tolerances) with the corrected causal oracle (oracle/attention_oracle.py; the reference's own
``reference_attention(causal=True)`` masks the wrong axes).  Everything goes through the public wrappers
-> flashattention_lab_cuda shim -> C ABI -> sm_100a kernels; the oracle runs on the CPU copy of the same inputs."""
import pytest
import torch

from fa3.cuda.impl import fa3_cuda
from fa3.op import fa3_attention
from fa3.spec import pick_fa3_spec
from oracle.attention_oracle import dense_backward, dense_forward
from tests.utils import LSE_TOL, assert_allclose, dtype_tolerances, flatten_lse, flatten_output, make_qkv

pytestmark = pytest.mark.gpu
EXTRA = (False,)  # trailing positional args of fa3_cuda after `spec`


def _oracle_fwd(q, k, v, causal, scale):
    return dense_forward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), causal=causal, softmax_scale=scale)


# reference tests/test_correctness_fa3.py "torch forward" shapes, run through the CUDA path instead
@pytest.mark.parametrize("shape", [(1, 2, 24, 32)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fa3_cuda_forward_matches_reference(shape, causal, merge_heads, dtype, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(20)
    batch, heads, seqlen, head_dim = shape
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    scale = head_dim ** -0.5
    o, lse = fa3_cuda(q, k, v, causal, scale, pick_fa3_spec(head_dim), *EXTRA)
    assert o.shape == q.shape and o.dtype == dtype and lse.dtype == torch.float32
    assert lse.shape == q.shape[:-1]
    o_ref, lse_ref = _oracle_fwd(q, k, v, causal, scale)
    assert_allclose(flatten_output(o), flatten_output(o_ref), **dtype_tolerances(dtype))
    assert_allclose(flatten_lse(lse), flatten_lse(lse_ref), **LSE_TOL)


# reference tests/test_correctness_fa3.py::test_fa3_cuda_backward_matches_reference
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("merge_heads", [True, False])
def test_fa3_cuda_backward_matches_reference(causal, merge_heads, cuda_extension_available):
    assert cuda_extension_available
    torch.manual_seed(22)
    dtype = torch.float16
    batch, heads, seqlen, head_dim = 1, 2, 32, 32
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=dtype, merge_heads=merge_heads)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, _ = fa3_cuda(q, k, v, causal, scale, pick_fa3_spec(head_dim), *EXTRA)
    do = torch.randn_like(o)
    (o * do).sum().backward()
    dq_ref, dk_ref, dv_ref, _, _ = dense_backward(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(),
                                                  causal, scale)
    tol = dtype_tolerances(dtype)
    assert_allclose(q.grad, dq_ref, **tol)
    assert_allclose(k.grad, dk_ref, **tol)
    assert_allclose(v.grad, dv_ref, **tol)


# the extra fp32 "torch backward" shape of the reference suite (head dim 64), as 16-bit through the CUDA path
@pytest.mark.parametrize("causal", [False, True])
def test_fa3_cuda_backward_odd_head_dim(causal, cuda_extension_available):
    torch.manual_seed(21)
    batch, heads, seqlen, head_dim = 1, 2, 64, 64
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.bfloat16, merge_heads=True)
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    scale = head_dim ** -0.5
    o, lse = fa3_cuda(q, k, v, causal, scale, pick_fa3_spec(head_dim), *EXTRA)
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


# reference test_fa3_backend_consistency: shape/seed kept; the only other "backend" left is the oracle
def test_fa3_attention_entry_point_consistency(cuda_extension_available):
    torch.manual_seed(23)
    batch, heads, seqlen, head_dim = 1, 2, 20, 32
    q, k, v = make_qkv(batch, heads, seqlen, head_dim, device="cuda", dtype=torch.float16, merge_heads=False)
    o_auto, lse_auto = fa3_attention(q, k, v, causal=True)  # default scale d**-0.5, backend="auto"
    o_cuda, lse_cuda = fa3_attention(q, k, v, causal=True, softmax_scale=head_dim ** -0.5, backend="cuda")
    assert o_auto.shape == (batch, heads, seqlen, head_dim) and lse_auto.shape == (batch, heads, seqlen)
    assert torch.equal(o_auto, o_cuda) and torch.equal(lse_auto, lse_cuda)
    o_ref, lse_ref = _oracle_fwd(q, k, v, True, head_dim ** -0.5)
    assert_allclose(flatten_output(o_auto), flatten_output(o_ref), **dtype_tolerances(torch.float16))
    assert_allclose(flatten_lse(lse_auto), flatten_lse(lse_ref), **LSE_TOL)


def test_fa3_fp8_flag_is_rejected_loudly(cuda_extension_available):
    """fp8=True stays in the signature (reference src/fa3/op.py:7) but the reference's emulation is broken (D5) and
    un-pinned, so it raises instead of silently returning something else."""
    q, k, v = make_qkv(1, 2, 32, 32, device="cuda", dtype=torch.float16, merge_heads=True)
    with pytest.raises(NotImplementedError):
        fa3_cuda(q, k, v, False, 32 ** -0.5, pick_fa3_spec(32), True)
    with pytest.raises(NotImplementedError):
        fa3_attention(q, k, v, fp8=True)
