"""CPU: host logic of the multi-GPU drivers.  The ring schedule is exercised (a) in loopback for P in {1,2,4} and
(b) for real with torch.distributed/gloo at world_size 2, with a CPU block operator built on the oracle (test-only:
the product default is the CUDA library).  Expected results come from single-device dense attention."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from dist.ring import (BlockOps, contiguous_split, ring_attention, ring_backward, ring_forward, run_loopback,
                       zigzag_chunk_ids, zigzag_merge, zigzag_split)
from dist.shard import shard_range, sharded_attention
from oracle.attention_oracle import blocked_backward, dense_backward_fp32, dense_forward, merge_partials


# ---------------------------------------------------------------------------------------------------------------------
# a CPU BlockOps with the semantics of the shim's raw entry points (fp32 tensors)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_fwd(q, k, v, causal, scale, *, q_row0=0, kv_col0=0, out=None, lse=None, merge=False):
    o_new, lse_new = dense_forward(q, k, v, causal, scale, q_row0, kv_col0)
    if merge:
        o_m, lse_m = merge_partials(out, lse, o_new, lse_new)
        out.copy_(o_m.to(out.dtype))
        lse.copy_(lse_m)
    else:
        out.copy_(o_new)
        lse.copy_(lse_new)
    return out, lse


def _cpu_prepare(o, do, lse, zero=None):
    if zero is not None:
        zero.zero_()
    return {"delta_o": o.clone(), "lse": lse.clone()}


def _cpu_bwd(q, k, v, o, do, lse, causal, scale, *, q_row0=0, kv_col0=0, rowstats=None, dq_accum=None,
             dk_accum=None, dv_accum=None, accum_overwrite=False):
    dq, dk, dv = blocked_backward(q, k, v, rowstats["delta_o"], do, rowstats["lse"], causal, scale, 16, 16, q_row0,
                                  kv_col0, out_dtype=torch.float32)
    dq_accum += dq / scale  # the CUDA kernel accumulates UNSCALED dQ partials; finish() applies the scale
    if dk_accum is None:
        return None, dk, dv
    if accum_overwrite:     # ring form: fp32 partials, stored ...
        dk_accum.copy_(dk)
        dv_accum.copy_(dv)
    else:                   # ... or added
        dk_accum += dk
        dv_accum += dv
    return None, None, None


def _cpu_finish(dq_accum, dtype, scale):
    return (dq_accum * scale).to(dtype)


CPU_OPS = BlockOps(fwd=_cpu_fwd, prepare=_cpu_prepare, bwd=_cpu_bwd, finish=_cpu_finish)


def test_shard_ranges_tile_exactly():
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_sharded_attention_equals_global():
    torch.manual_seed(0)
    q, k, v = (torch.randn(6, 20, 16) for _ in range(3))
    o_ref, lse_ref = dense_forward(q, k, v, True)
    for world in (2, 4):
        for r in range(world):
            (b, e), (o, lse) = sharded_attention(lambda a, b_, c, **kw: dense_forward(a, b_, c, **kw), q, k, v, r, world,
                                                 causal=True)
            torch.testing.assert_close(o, o_ref[b:e])
            torch.testing.assert_close(lse, lse_ref[b:e])


def test_zigzag_roundtrip_and_balance():
    x = torch.arange(2 * 16 * 3.0).reshape(2, 16, 3)
    for world in (1, 2, 4):
        parts = zigzag_split(x, world)
        assert torch.equal(zigzag_merge(parts), x)
        assert all(p.shape[1] == 16 // world for p in parts)
        # causal work per rank (number of visible (q,k) chunk pairs) is identical for every rank
        work = []
        for r in range(world):
            mine = zigzag_chunk_ids(r, world)
            work.append(sum(1 + qc for qc in mine))
        assert len(set(work)) == 1
    with pytest.raises(ValueError):
        zigzag_split(x[:, :15], 2)


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("causal", [False, True])
def test_ring_loopback_matches_single_device(world, causal):
    torch.manual_seed(1)
    bh, n, d = 2, 32 * world, 16
    q, k, v, do = (torch.randn(bh, n, d) for _ in range(4))
    scale = d ** -0.5
    dq_ref, dk_ref, dv_ref, o_ref, lse_ref = dense_backward_fp32(q, k, v, do, causal, scale)
    split = zigzag_split if causal else contiguous_split
    merge = zigzag_merge if causal else (lambda parts: torch.cat(parts, dim=-2))
    ql, kl, vl, dol = (split(t, world) for t in (q, k, v, do))
    outs = run_loopback([ring_forward(CPU_OPS, r, world, ql[r], kl[r], vl[r], causal, scale) for r in range(world)])
    o = merge([x[0] for x in outs])
    lse = merge([x[1].unsqueeze(-1) for x in outs]).squeeze(-1)
    torch.testing.assert_close(o, o_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lse, lse_ref, rtol=1e-4, atol=1e-5)
    grads = run_loopback([ring_backward(CPU_OPS, r, world, ql[r], kl[r], vl[r], outs[r][0], outs[r][1], dol[r], causal,
                                        scale) for r in range(world)])
    for idx, ref in enumerate((dq_ref, dk_ref, dv_ref)):
        torch.testing.assert_close(merge([g[idx] for g in grads]), ref, rtol=1e-3, atol=1e-4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, causal, q, k, v, do, out_queue):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        split = zigzag_split if causal else contiguous_split
        ql, kl, vl, dol = (split(t, world)[rank].requires_grad_(t is not do) for t in (q, k, v, do))
        o, lse = ring_attention(ql, kl, vl, causal=causal, ops=CPU_OPS)
        o.backward(dol)
        # plain numpy (pickled by value): torch tensors travel as shared-memory handles that vanish when this process exits
        out_queue.put((rank,) + tuple(t.detach().numpy().copy() for t in (o, lse, ql.grad, kl.grad, vl.grad)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,causal", [(2, False), (2, True), (3, True), (4, False)])
def test_ring_attention_gloo(world, causal):
    # world 2 has prev == next; 3 and 4 exercise distinct ring neighbours in the send/recv driver
    torch.manual_seed(2)
    q, k, v, do = (torch.randn(2, 24 * world, 16) for _ in range(4))
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, causal, q, k, v, do, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((queue.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    got = [(g[0],) + tuple(torch.from_numpy(a) for a in g[1:]) for g in got]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merge = zigzag_merge if causal else (lambda parts: torch.cat(parts, dim=-2))
    dq_ref, dk_ref, dv_ref, o_ref, lse_ref = dense_backward_fp32(q, k, v, do, causal)
    torch.testing.assert_close(merge([g[1] for g in got]), o_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(merge([g[2].unsqueeze(-1) for g in got]).squeeze(-1), lse_ref, rtol=1e-4, atol=1e-5)
    for idx, ref in ((3, dq_ref), (4, dk_ref), (5, dv_ref)):
        torch.testing.assert_close(merge([g[idx] for g in got]), ref, rtol=1e-3, atol=1e-4)
