"""GPU: block-sparse tile mask and dropout through the C ABI (fa_sm100_fwd_ex / bwd_ex) against
the oracle's fp32 restatement, which regenerates the kernels' Philox keep mask bit for bit on the CPU."""
import pytest
import torch

import flashattention_lab_cuda as ext
from fa3 import fa3_block_sparse_attention
from oracle.attention_oracle import dense_ext_backward_fp32, dropout_keep_mask, error_report

pytestmark = pytest.mark.gpu


def _check(q, k, v, do, causal, scale, mask, p, seed, tol=5e-2):
    o, lse = ext.fwd_ex_raw(q, k, v, causal, scale, block_mask=mask, dropout_p=p, seed=seed)
    dq, dk, dv = ext.bwd_ex_raw(q, k, v, o, do, lse, causal, scale, block_mask=mask, dropout_p=p, seed=seed)
    dq_r, dk_r, dv_r, o_r, lse_r = dense_ext_backward_fp32(q, k, v, do, causal, scale, mask, p, seed)
    for name, got, want, t in (("o", o, o_r, tol), ("lse", lse, lse_r, 1e-3), ("dq", dq, dq_r, tol), ("dk", dk, dk_r, tol),
                               ("dv", dv, dv_r, tol)):
        rep = error_report(got, want, t, t)
        assert rep["violations"] == 0, f"{name}: {rep}"
    return o, lse, dq, dk, dv


def _random_mask(shape, density, seed, keep_diagonal=True):
    g = torch.Generator().manual_seed(seed)
    m = (torch.rand(shape, generator=g) < density).to(torch.int32)
    if keep_diagonal:  # every query tile sees at least its own diagonal tile (no all-masked rows under causal)
        idx = torch.arange(min(shape[-2:]))
        m[..., idx, idx] = 1
    return m


@pytest.mark.parametrize("bh,n,d,dtype,causal", [(2, 1024, 128, torch.bfloat16, True), (3, 777, 64, torch.float16, False),
                                                 (2, 1300, 96, torch.bfloat16, True), (1, 2048, 128, torch.float16, False)])
@pytest.mark.parametrize("per_slice", [False, True])
def test_block_sparse_mask(bh, n, d, dtype, causal, per_slice):
    torch.manual_seed(n)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(4))
    nb = (n + 127) // 128
    mask = _random_mask((bh, nb, nb) if per_slice else (nb, nb), 0.4, seed=n + d).cuda()
    _check(q, k, v, do, causal, d ** -0.5, mask, 0.0, 0)


def test_masked_out_rows_and_columns():
    """A query tile with no active tile yields O = 0, lse = -inf and zero gradients; a K/V tile nobody looks at gets
    zero dK/dV (its CTA has no iterations at all)."""
    torch.manual_seed(1)
    bh, n, d = 2, 640, 128
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    mask = torch.ones(5, 5, dtype=torch.uint8)
    mask[2, :] = 0
    mask[:, 3] = 0
    o, lse, dq, dk, dv = _check(q, k, v, do, False, d ** -0.5, mask.cuda(), 0.0, 0)
    assert torch.all(o[:, 256:384] == 0) and torch.all(torch.isinf(lse[:, 256:384]))
    assert torch.all(dq[:, 256:384] == 0) and torch.all(dk[:, 384:512] == 0) and torch.all(dv[:, 384:512] == 0)


def test_full_mask_without_dropout_is_the_dense_kernel():
    torch.manual_seed(2)
    q, k, v, do = (torch.randn(2, 900, 128, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    o, lse = ext.fwd_raw(q, k, v, True, 0.09)
    nb = (900 + 127) // 128
    o2, lse2 = ext.fwd_ex_raw(q, k, v, True, 0.09, block_mask=torch.ones(nb, nb, device="cuda"))
    assert torch.equal(o, o2) and torch.equal(lse, lse2)
    o3, lse3 = ext.fwd_ex_raw(q, k, v, True, 0.09)  # no extras at all: dispatches to the dense kernel
    assert torch.equal(o, o3) and torch.equal(lse, lse3)
    dk_a = ext.bwd_raw(q, k, v, o, do, lse, True, 0.09)[1]
    dk_b = ext.bwd_ex_raw(q, k, v, o, do, lse, True, 0.09, block_mask=torch.ones(nb, nb, device="cuda"))[1]
    assert torch.equal(dk_a, dk_b)


@pytest.mark.parametrize("bh,n,d,dtype,causal,p", [(2, 512, 128, torch.bfloat16, True, 0.1), (2, 700, 64, torch.float16, False, 0.5),
                                                   (1, 1536, 128, torch.bfloat16, True, 0.25)])
def test_dropout(bh, n, d, dtype, causal, p):
    torch.manual_seed(n)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(4))
    o, lse, *_ = _check(q, k, v, do, causal, d ** -0.5, None, p, seed=4242 + n, tol=6e-2)
    # lse ignores dropout; a different seed gives a different output, the same seed the same one (bit for bit)
    o_dense, lse_dense = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    assert torch.equal(lse, lse_dense)
    o_again, _ = ext.fwd_ex_raw(q, k, v, causal, d ** -0.5, dropout_p=p, seed=4242 + n)
    o_other, _ = ext.fwd_ex_raw(q, k, v, causal, d ** -0.5, dropout_p=p, seed=1)
    assert torch.equal(o, o_again) and not torch.equal(o, o_other)


def test_dropout_keep_bits_match_the_oracle_exactly():
    """V = identity columns turns O into the (dropped, rescaled) probabilities themselves: wherever the oracle's keep mask
    is False the kernel's output must be exactly 0, and nowhere else (for probabilities that are not tiny)."""
    torch.manual_seed(3)
    bh, n, d, p, seed = 2, 128, 128, 0.3, 777
    q, k = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) * 0.3 for _ in range(2))
    v = torch.eye(n, d, device="cuda", dtype=torch.bfloat16).expand(bh, n, d).contiguous()
    o, _ = ext.fwd_ex_raw(q, k, v, False, d ** -0.5, dropout_p=p, seed=seed)
    keep, scale = dropout_keep_mask(bh, n, n, p, seed)
    assert torch.equal((o.cpu() != 0), keep)
    probs = torch.softmax((q.float() @ k.float().transpose(-2, -1)) * d ** -0.5, dim=-1).cpu()
    assert (o.float().cpu() - probs * keep * scale).abs().max() < 2e-3


def test_block_sparse_with_dropout_and_sharding_offsets():
    torch.manual_seed(4)
    bh, n, d = 2, 1024, 128
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    mask = _random_mask((8, 8), 0.5, seed=5).cuda()
    o, lse, *_ = _check(q, k, v, do, True, d ** -0.5, mask, 0.2, seed=31)
    # the lower half of the queries as a separate call with global offsets sees the same random bits
    o_half, lse_half = ext.fwd_ex_raw(q[:, 512:].contiguous(), k, v, True, d ** -0.5, block_mask=mask[4:].contiguous(),
                                      dropout_p=0.2, seed=31, q_row0=512)
    assert torch.equal(o_half, o[:, 512:]) and torch.equal(lse_half, lse[:, 512:])


def test_public_entry_point_with_autograd():
    torch.manual_seed(5)
    b, h, n, d = 2, 3, 384, 64
    q, k, v = (torch.randn(b, h, n, d, device="cuda", dtype=torch.float16).requires_grad_(True) for _ in range(3))
    do = torch.randn(b, h, n, d, device="cuda", dtype=torch.float16)
    mask = torch.tensor([[1, 0, 0], [1, 1, 0], [0, 1, 1]])
    o, lse = fa3_block_sparse_attention(q, k, v, block_sparse_mask=mask, dropout_p=0.1, causal=True, seed=9)
    assert o.shape == (b, h, n, d) and lse.shape == (b, h, n)
    o.backward(do)
    dq_r, dk_r, dv_r, o_r, lse_r = dense_ext_backward_fp32(q.detach().reshape(-1, n, d), k.detach().reshape(-1, n, d),
                                                           v.detach().reshape(-1, n, d), do.reshape(-1, n, d), True,
                                                           None, mask, 0.1, 9)
    for got, want in ((o, o_r), (q.grad, dq_r), (k.grad, dk_r), (v.grad, dv_r)):
        assert error_report(got.reshape(-1, n, d), want, 5e-2, 5e-2)["violations"] == 0
    o_eval, _ = fa3_block_sparse_attention(q, k, v, block_sparse_mask=mask, dropout_p=0.1, causal=True, training=False)
    o_nodrop, _ = fa3_block_sparse_attention(q, k, v, block_sparse_mask=mask, causal=True)
    assert torch.equal(o_eval, o_nodrop)
    with pytest.raises(RuntimeError, match="block_sparse_mask has shape"):
        fa3_block_sparse_attention(q, k, v, block_sparse_mask=torch.ones(2, 2))
    with pytest.raises(ValueError):
        fa3_block_sparse_attention(q, k, v, dropout_p=1.0)
