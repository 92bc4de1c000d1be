"""GPU: ring attention (zig-zag causal + non-causal) with the sm_100a block operator, all P virtual ranks emulated in one
process on one GPU (the same schedule code the NCCL driver runs), against the CPU oracle on the gathered tensors."""
import pytest
import torch

from dist.ring import (contiguous_split, cuda_block_ops, ring_backward, ring_forward, run_loopback, zigzag_merge,
                       zigzag_split)
from oracle.attention_oracle import dense_backward_fp32, error_report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n", [(2, 1024), (4, 2048), (4, 1536)])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("dtype,d", [(torch.bfloat16, 128), (torch.float16, 64)])
def test_ring_loopback_cuda_matches_oracle(world, n, causal, dtype, d):
    torch.manual_seed(7)
    bh = 2
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(4))
    scale = d ** -0.5
    ops = cuda_block_ops()
    split = zigzag_split if causal else contiguous_split
    merge = zigzag_merge if causal else (lambda parts: torch.cat(parts, dim=-2))
    ql, kl, vl, dol = (split(t, world) for t in (q, k, v, do))
    outs = run_loopback([ring_forward(ops, r, world, ql[r], kl[r], vl[r], causal, scale) for r in range(world)])
    grads = run_loopback([ring_backward(ops, r, world, ql[r], kl[r], vl[r], outs[r][0], outs[r][1], dol[r], causal, scale)
                          for r in range(world)])
    torch.cuda.synchronize()
    o = merge([x[0] for x in outs])
    lse = merge([x[1].unsqueeze(-1) for x in outs]).squeeze(-1)
    dq_ref, dk_ref, dv_ref, o_ref, lse_ref = dense_backward_fp32(q.cpu(), k.cpu(), v.cpu(), do.cpu(), causal, scale)
    checks = [("o", o, o_ref, 5e-2), ("lse", lse, lse_ref, 1e-3)]
    checks += [(name, merge([g[i] for g in grads]), ref, 5e-2) for i, (name, ref) in
               enumerate((("dq", dq_ref), ("dk", dk_ref), ("dv", dv_ref)))]
    for name, got, want, tol in checks:
        rep = error_report(got, want, tol, tol)
        assert rep["violations"] == 0, f"{name}: {rep}"


def test_offsets_against_oracle_directly():
    """The two generalisations the ring relies on: global row/column offsets in the causal mask and n_q != n_kv."""
    import flashattention_lab_cuda as ext
    from oracle.attention_oracle import dense_backward_fp32 as oracle_bwd

    torch.manual_seed(8)
    bh, n_q, n_kv, d = 3, 200, 456, 128
    q, do = (torch.randn(bh, n_q, d, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    k, v = (torch.randn(bh, n_kv, d, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    for q_row0, kv_col0 in ((300, 0), (256, 128), (0, 0)):
        o, lse = ext.fwd_raw(q, k, v, True, 0.09, q_row0=q_row0, kv_col0=kv_col0)
        dq, dk, dv = ext.bwd_raw(q, k, v, o, do, lse, True, 0.09, q_row0=q_row0, kv_col0=kv_col0)
        dq_r, dk_r, dv_r, o_r, lse_r = oracle_bwd(q.cpu(), k.cpu(), v.cpu(), do.cpu(), True, 0.09, q_row0, kv_col0)
        for name, got, want, tol in (("o", o, o_r, 5e-2), ("lse", lse, lse_r, 1e-3), ("dq", dq, dq_r, 5e-2),
                                     ("dk", dk, dk_r, 5e-2), ("dv", dv, dv_r, 5e-2)):
            rep = error_report(got, want, tol, tol)
            assert rep["violations"] == 0, f"offsets {(q_row0, kv_col0)} {name}: {rep}"


def test_rows_without_visible_keys():
    import flashattention_lab_cuda as ext

    torch.manual_seed(9)
    q, k, v = (torch.randn(2, 256, 64, device="cuda", dtype=torch.float16) for _ in range(3))
    o, lse = ext.fwd_raw(q, k, v, True, 0.125, q_row0=0, kv_col0=100)  # rows 0..99 see nothing
    assert torch.all(o[:, :100] == 0) and torch.all(torch.isinf(lse[:, :100]) & (lse[:, :100] < 0))
    assert torch.isfinite(lse[:, 100:]).all() and torch.isfinite(o.float()).all()
