"""Ring attention over a sequence-sharded context (BASELINE config C5).

The reference has no distributed code; the exchange pattern is derived from its online-softmax update
(reference ``src/fa1/torch/impl.py:53-62``): attention over a union of key blocks is the log-sum-exp merge of the
per-block partials.  One process per GPU; K/V blocks (and, in the backward, their fp32 dK/dV accumulators) travel
around the ring with NCCL send/recv while the current block is being processed.

Layout (causal): the sequence is cut into 2P chunks of c rows; rank r owns chunks r and 2P-1-r ("zig-zag"), stored
as one local tensor (bh, 2c, d) = [chunk r | chunk 2P-1-r].  Every ring step is then exactly one kernel launch of equal
cost on every rank:

    source rank p == r :  [chunk_a | chunk_b] x [K_a | K_b]  causal in LOCAL indices (one launch)
    p <  r             :  [chunk_a | chunk_b] x K_a          (all visible, non-causal call)
    p >  r             :  chunk_b x [K_a | K_b]              (all visible, non-causal call)

(the local [chunk_a | chunk_b] tensor behaves like a contiguous sequence under a causal mask: every key of chunk a
precedes chunk b, the rows between them belong to other ranks, and K_b is chunk b's own diagonal block).  Partials are
merged inside the forward kernel's epilogue (``merge=True``).  Non-causal attention needs no zig-zag: every step is
[all local q] x [visiting block].

Backward: dQ accumulates locally in fp32.  The dK/dV of a K/V block accumulate in fp32 buffers that TRAVEL with the
block.  Each step is one launch that writes this rank's partial in fp32 (``dk_accum``/``dv_accum`` in overwrite mode:
no 16-bit rounding); the accumulators, which cross NVLink while that launch runs, take the partial with one fp32 add
and move on.  (Measured alternatives, profiles/r02i: letting the kernel reduce-add straight into the travelling
buffers needs them to have arrived before the launch; splitting every block into two key halves with their own launches
buys that slack but doubles the launches -- 6576 vs 7584 TFLOP/s at 8 GPUs for the one-launch schedule.)

The schedule is written once as a generator that yields its communication requests, so the same code runs
  * under ``torch.distributed`` (``SymmMemRingDriver``: copy-engine peer copies out of symmetric memory over NVLink, the
    default on GPUs; ``TorchRingDriver``: NCCL/gloo batched isend/irecv), both overlapped with compute, and
  * inside ONE process for all P ranks (``run_loopback``) — used by the single-GPU and CPU tests.
The block operator is injected (``BlockOps``); the product default is the sm_100a library.  Tests may inject a CPU
implementation — nothing in this module imports the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Generator, List, Optional, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------------------------------------------------
# zig-zag partition helpers
# ----------------------------------------------------------------------------------------------------------------------


def zigzag_chunk_ids(rank: int, world: int) -> Tuple[int, int]:
    return rank, 2 * world - 1 - rank


def zigzag_split(x: torch.Tensor, world: int, dim: int = -2) -> List[torch.Tensor]:
    """Global tensor -> per-rank local tensors [chunk r | chunk 2P-1-r] along ``dim``."""
    n = x.shape[dim]
    if n % (2 * world):
        raise ValueError(f"sequence length {n} must be a multiple of 2*world = {2 * world}")
    chunks = x.chunk(2 * world, dim=dim)
    return [torch.cat([chunks[a], chunks[b]], dim=dim).contiguous()
            for a, b in (zigzag_chunk_ids(r, world) for r in range(world))]


def zigzag_merge(parts: Sequence[torch.Tensor], dim: int = -2) -> torch.Tensor:
    """Inverse of ``zigzag_split``."""
    world = len(parts)
    chunks: List[Optional[torch.Tensor]] = [None] * (2 * world)
    for r, p in enumerate(parts):
        a, b = p.chunk(2, dim=dim)
        ia, ib = zigzag_chunk_ids(r, world)
        chunks[ia], chunks[ib] = a, b
    return torch.cat(chunks, dim=dim)


def contiguous_split(x: torch.Tensor, world: int, dim: int = -2) -> List[torch.Tensor]:
    return [c.contiguous() for c in x.chunk(world, dim=dim)]


# ----------------------------------------------------------------------------------------------------------------------
# block operator
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class BlockOps:
    """The single-device attention primitives the ring is built from (signatures of the shim's raw entry points)."""

    fwd: Callable      # (q, k, v, causal, scale, *, q_row0, kv_col0, out, lse, merge) -> (out, lse)
    prepare: Callable  # (o, do, lse) -> rowstats
    bwd: Callable      # (q, k, v, o, do, lse, causal, scale, *, q_row0, kv_col0, rowstats, dq_accum, dk_accum, dv_accum,
    #                    accum_overwrite): adds the fp32 dQ partial into dq_accum; dK/dV partials are added into (or,
    #                    with accum_overwrite, stored to) dk_accum / dv_accum
    finish: Callable   # (dq_accum, dtype, scale) -> dq


def cuda_block_ops() -> BlockOps:
    """The product operator: hand-written sm_100a kernels through the C ABI.  Raises if the library is missing."""
    import flashattention_lab_cuda as ext

    ext.load_library()
    return BlockOps(fwd=ext.fwd_raw, prepare=ext.bwd_prepare_raw, bwd=ext.bwd_raw, finish=ext.dq_finish_raw)


# ----------------------------------------------------------------------------------------------------------------------
# the schedule (generators: ``tag, payload = yield``-style coroutines driven by a communicator)
# ----------------------------------------------------------------------------------------------------------------------
# yield ("post", [tensors to send to rank+1])  -> handle        (receive buffers come from rank-1)
# yield ("wait", handle)                       -> [tensors received from rank-1]
Coroutine = Generator[tuple, object, tuple]


def _check_local(q, k, v, causal):
    if q.dim() != 3 or k.shape != v.shape or k.shape[0] != q.shape[0] or k.shape[2] != q.shape[2]:
        raise ValueError("ring attention: q, k, v must be (bh, n_local, d) with matching bh and d")
    if k.shape[1] != q.shape[1]:
        raise ValueError("ring attention: k/v must have q's n_local (every rank owns the same rows of q, k and v)")
    if q.shape[1] % 2:
        raise ValueError("ring attention: n_local must be even (zig-zag chunks / key halves)")


def ring_forward(ops: BlockOps, rank: int, world: int, q, k, v, causal: bool, scale: float) -> Coroutine:
    """q, k, v: local (bh, n_local, d).  Returns (o, lse) for the local rows."""
    _check_local(q, k, v, causal)
    bh, n_local, d = q.shape
    c = n_local // 2
    o = torch.empty_like(q)
    lse = torch.empty((bh, n_local), device=q.device, dtype=torch.float32)
    kv = [k, v]
    for step in range(world):
        handle = None
        if step + 1 < world:
            handle = yield ("post", kv)  # next block starts moving before this one is consumed
        src = (rank - step) % world
        kb, vb = kv
        if not causal:
            ops.fwd(q, kb, vb, False, scale, out=o, lse=lse, merge=step > 0)
        elif src == rank:  # own block: the local rows are causally ordered as they lie
            ops.fwd(q, kb, vb, True, scale, out=o, lse=lse, merge=False)
        elif src < rank:
            ops.fwd(q, kb[:, :c], vb[:, :c], False, scale, out=o, lse=lse, merge=True)
        else:
            ops.fwd(q[:, c:], kb, vb, False, scale, out=o[:, c:], lse=lse[:, c:], merge=True)
        if handle is not None:
            kv = yield ("wait", handle)
    return o, lse


def ring_backward(ops: BlockOps, rank: int, world: int, q, k, v, o, lse, do, causal: bool, scale: float) -> Coroutine:
    """Returns (dq, dk, dv) for the local rows.  dQ accumulates locally in fp32.  Every step is ONE launch that stores
    this rank's dK/dV partial of the visiting block in fp32 (no 16-bit rounding); the block's fp32 accumulators, which
    arrive from the previous rank while that launch runs, take the partial and move on, so each block is home again,
    complete, after ``world`` hops."""
    _check_local(q, k, v, causal)
    bh, n_local, d = q.shape
    c = n_local // 2
    dq_acc = torch.empty(q.shape, device=q.device, dtype=torch.float32)
    stats_all = ops.prepare(o, do, lse, zero=dq_acc)
    stats_b = ops.prepare(o[:, c:], do[:, c:], lse[:, c:]) if causal else None
    kv = [k, v]
    acc = None
    acc_handle = None
    for step in range(world):
        kv_handle = None
        if step + 1 < world:
            kv_handle = yield ("post", kv)
        src = (rank - step) % world
        kb, vb = kv
        rows = slice(0, c) if (causal and src < rank) else slice(None)  # key rows of the block this rank can see
        part = [torch.empty(kb[:, rows].shape, device=k.device, dtype=torch.float32) for _ in range(2)]
        kw = dict(dk_accum=part[0], dv_accum=part[1], accum_overwrite=True)
        if not causal:
            ops.bwd(q, kb, vb, None, do, None, False, scale, rowstats=stats_all, dq_accum=dq_acc, **kw)
        elif src == rank:  # own block: the local rows are causally ordered as they lie
            ops.bwd(q, kb, vb, None, do, None, True, scale, rowstats=stats_all, dq_accum=dq_acc, **kw)
        elif src < rank:   # every local row sees all of chunk a of an earlier rank; its chunk b is invisible
            ops.bwd(q, kb[:, :c], vb[:, :c], None, do, None, False, scale, rowstats=stats_all, dq_accum=dq_acc, **kw)
        else:              # only chunk b sees a later rank's keys, and it sees both of its chunks
            ops.bwd(q[:, c:], kb, vb, None, do[:, c:], None, False, scale, rowstats=stats_b, dq_accum=dq_acc[:, c:],
                    **kw)
        if acc_handle is None:
            acc = part  # step 0 works on the rank's own block: its accumulators start as this partial
        else:           # accumulators of the block we are working on arrive from the previous rank
            acc = yield ("wait", acc_handle)
            acc[0][:, rows] += part[0]
            acc[1][:, rows] += part[1]
        acc_handle = yield ("post", acc)  # they move on with their block (the last hop brings ours home)
        if kv_handle is not None:
            kv = yield ("wait", kv_handle)
    acc = yield ("wait", acc_handle)
    dq = ops.finish(dq_acc, q.dtype, scale)
    return dq, acc[0].to(k.dtype), acc[1].to(v.dtype)


# ----------------------------------------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------------------------------------
def run_loopback(coroutines: Sequence[Coroutine]) -> List[tuple]:
    """Run the P per-rank coroutines of one collective call inside one process (ranks execute in lock-step between
    communication points; sends are delivered to rank+1 by reference)."""
    world = len(coroutines)
    results: List[Optional[tuple]] = [None] * world
    mailbox = [dict() for _ in range(world)]  # mailbox[r][seq] = tensors posted by rank r
    seq = [0] * world
    pending = [None] * world  # value to send into each coroutine next
    live = list(range(world))
    # advance every coroutine to its first yield
    requests = {}
    for r in live:
        try:
            requests[r] = next(coroutines[r])
        except StopIteration as stop:
            results[r] = stop.value
    while requests:
        progressed = False
        for r in sorted(requests):
            kind, payload = requests[r]
            if kind == "post":
                mailbox[r][seq[r]] = [t for t in payload]
                reply = (r, seq[r])
                seq[r] += 1
            else:  # wait on a handle (owner rank, seq): data comes from rank-1's post with the same sequence number
                _, s = payload
                src = (r - 1) % world
                if s not in mailbox[src]:
                    continue  # neighbour has not posted yet: let the others run
                reply = [t.clone() for t in mailbox[src].pop(s)]
            progressed = True
            try:
                requests[r] = coroutines[r].send(reply)
            except StopIteration as stop:
                results[r] = stop.value
                del requests[r]
        if not progressed:
            raise RuntimeError("ring loopback deadlock: every rank is waiting")
    return results  # type: ignore[return-value]


class TorchRingDriver:
    """Drives one rank's coroutine with torch.distributed point-to-point ops (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.next_rank = dist.get_global_rank(group, (self.rank + 1) % self.world) if group else (self.rank + 1) % self.world
        self.prev_rank = dist.get_global_rank(group, (self.rank - 1) % self.world) if group else (self.rank - 1) % self.world

    def _post(self, tensors):
        dist = self.dist
        sends = [t.contiguous() for t in tensors]
        recvs = [torch.empty_like(t) for t in sends]
        ops = []
        for s, r in zip(sends, recvs):
            ops.append(dist.P2POp(dist.isend, s, self.next_rank, self.group))
            ops.append(dist.P2POp(dist.irecv, r, self.prev_rank, self.group))
        reqs = dist.batch_isend_irecv(ops)
        return reqs, recvs, sends  # keep `sends` alive until the wait

    def run(self, coroutine: Coroutine) -> tuple:
        try:
            request = next(coroutine)
            while True:
                kind, payload = request
                if kind == "post":
                    reply = self._post(payload)
                else:
                    reqs, recvs, _sends = payload
                    for rq in reqs:
                        rq.wait()
                    reply = recvs
                request = coroutine.send(reply)
        except StopIteration as stop:
            return stop.value


class SymmMemRingDriver:
    """Same coroutine protocol, but blocks move by COPY-ENGINE peer copies out of symmetric (peer-mapped) memory over
    NVLink instead of NCCL's SM-resident send/recv kernels.  The attention kernels own every SM (one 200 KiB CTA each),
    so NCCL's copy CTAs are starved until a kernel's tail and the exchange stops overlapping with compute once the
    per-step compute gets short (measured: C5 at 8 GPUs is comm-bound at an effective ~70 GB/s with NCCL P2P, while an
    idle-GPU NCCL exchange does 313 GiB/s and a copy-engine peer pull 675 GiB/s on the same box).

    post(tensors): stage them into this rank's symmetric slot (HBM copy on the compute stream); then, on a side stream:
    cross-rank barrier (every slot staged) -> pull the previous rank's slot into fresh local tensors (DMA).
    wait(handle):  the compute stream waits for that pull.  Two slots per message signature; slot reuse is safe because
    a rank reaches barrier n+1 only after its pull n, and post n+2 is issued after wait n+1."""

    _drivers = {}          # keyed on the process-group OBJECT (an id() can be recycled after the group is collected)
    MAX_CHANNELS = 8       # message signatures kept per driver; older symmetric slots are dropped (LRU)

    @classmethod
    def get(cls, group=None):
        import torch.distributed as dist

        key = group or dist.group.WORLD
        if key not in cls._drivers:
            cls._drivers[key] = cls(group)
        return cls._drivers[key]

    @classmethod
    def close_all(cls):
        """Drop every cached driver and its symmetric slots (call before destroying the process groups)."""
        for drv in cls._drivers.values():
            drv.channels.clear()
        cls._drivers.clear()

    @staticmethod
    def usable(example: torch.Tensor, group=None) -> bool:
        """Symmetric-memory peer copies need a single-node group whose ring neighbours are P2P-mapped."""
        import torch.distributed as dist

        if not example.is_cuda or not dist.is_initialized() or dist.get_backend(group) != "nccl":
            return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except ImportError:
            return False
        world = dist.get_world_size(group)
        if world > torch.cuda.device_count():  # more ranks than local devices: the group spans nodes
            return False
        me = example.device.index if example.device.index is not None else torch.cuda.current_device()
        prev_global = dist.get_global_rank(group, (dist.get_rank(group) - 1) % world) if group else (dist.get_rank() - 1) % world
        peer = prev_global % torch.cuda.device_count()
        return peer == me or torch.cuda.can_device_access_peer(me, peer)

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group or dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.prev = (self.rank - 1) % self.world
        self.side = torch.cuda.Stream()
        self.channels = {}

    def _channel(self, tensors):
        import torch.distributed._symmetric_memory as symm

        sig = tuple((tuple(t.shape), t.dtype) for t in tensors)
        ch = self.channels.pop(sig, None)
        if ch is not None:
            self.channels[sig] = ch  # most recently used last
        if ch is None:
            while len(self.channels) >= self.MAX_CHANNELS:
                self.channels.pop(next(iter(self.channels)))  # least recently used symmetric slot pair
            offs, total = [], 0
            for t in tensors:
                offs.append(total)
                total += (t.numel() * t.element_size() + 255) // 256 * 256
            buf = symm.empty(2 * total, dtype=torch.uint8, device=tensors[0].device)
            hdl = symm.rendezvous(buf, self.group)
            ch = {"hdl": hdl, "buf": buf, "offs": offs, "slot_bytes": total, "count": 0}
            self.channels[sig] = ch
        return ch

    @staticmethod
    def _views(ch, rank, slot, tensors):
        return [ch["hdl"].get_buffer(rank, tuple(t.shape), t.dtype, (slot * ch["slot_bytes"] + off) // t.element_size())
                for t, off in zip(tensors, ch["offs"])]

    def _post(self, tensors):
        tensors = [t.contiguous() for t in tensors]
        ch = self._channel(tensors)
        slot = ch["count"] & 1
        ch["count"] += 1
        main = torch.cuda.current_stream()
        for dst, src in zip(self._views(ch, self.rank, slot, tensors), tensors):
            dst.copy_(src)
        staged = torch.cuda.Event()
        staged.record(main)
        recvs = [torch.empty_like(t) for t in tensors]
        with torch.cuda.stream(self.side):
            self.side.wait_event(staged)
            ch["hdl"].barrier(channel=0)
            for dst, src in zip(recvs, self._views(ch, self.prev, slot, tensors)):
                dst.copy_(src)
            done = torch.cuda.Event()
            done.record(self.side)
        for r in recvs:
            r.record_stream(self.side)
        return done, recvs

    def run(self, coroutine: Coroutine) -> tuple:
        try:
            request = next(coroutine)
            while True:
                kind, payload = request
                if kind == "post":
                    reply = self._post(payload)
                else:
                    done, recvs = payload
                    torch.cuda.current_stream().wait_event(done)
                    reply = recvs
                request = coroutine.send(reply)
        except StopIteration as stop:
            return stop.value


def _pick_transport(example: torch.Tensor, group=None, transport: str = "auto") -> str:
    import os

    transport = os.environ.get("FA_RING_TRANSPORT", transport)
    if transport == "auto":  # symmetric memory only where it can work; everything else takes the send/recv driver
        transport = "symm" if SymmMemRingDriver.usable(example, group) else "nccl"
    return transport


def make_driver(example: torch.Tensor, group=None, transport: str = "auto"):
    """'symm' = copy-engine peer copies over symmetric memory (single node, P2P-mapped CUDA peers),
    'nccl' = batched isend/irecv (NCCL on GPUs, gloo on CPU); 'auto' picks 'symm' only where it is usable."""
    if _pick_transport(example, group, transport) == "symm":
        return SymmMemRingDriver.get(group)
    return TorchRingDriver(group)


def ring_transport_name(example: torch.Tensor, group=None) -> str:
    return {"symm": "symmetric-memory peer pulls on the copy engines (NVLink)",
            "nccl": "NCCL P2P send/recv (NVLink)"}[_pick_transport(example, group)]


# ----------------------------------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------------------------------
class _RingAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, softmax_scale, group, ops):
        driver = make_driver(q, group)
        ops = ops or cuda_block_ops()
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = driver.run(ring_forward(ops, driver.rank, driver.world, q, k, v, bool(causal), float(softmax_scale)))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (bool(causal), float(softmax_scale), group, ops)
        return o, lse

    @staticmethod
    def backward(ctx, do, dlse):
        q, k, v, o, lse = ctx.saved_tensors
        causal, scale, group, ops = ctx.meta
        driver = make_driver(q, group)
        dq, dk, dv = driver.run(ring_backward(ops, driver.rank, driver.world, q, k, v, o, lse, do.contiguous(), causal,
                                              scale))
        return dq, dk, dv, None, None, None, None


_checked_shapes = set()


def _check_equal_across_ranks(q, group=None):
    """Every rank must own the same number of rows (the schedule slices received blocks at the LOCAL chunk size).
    Checked once per (group, shape) with one small all_gather — not per call, because under NCCL that collective's
    kernel would queue behind the SM-filling attention kernels of every step.  The cache assumes what the check then
    establishes: all ranks of the group call with the same sequence of shapes."""
    import torch.distributed as dist

    if not dist.is_initialized():
        return
    key = (group, tuple(q.shape[-2:]))
    if key in _checked_shapes:
        return
    world = dist.get_world_size(group)
    mine = torch.tensor([q.shape[-2], q.shape[-1]], device=q.device, dtype=torch.int64)
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine, group=group)
    if any(not torch.equal(t, mine) for t in everyone):
        raise ValueError(f"ring attention: ranks disagree on (n_local, d): {[t.tolist() for t in everyone]}")
    _checked_shapes.add(key)


def ring_attention(q, k, v, causal=False, softmax_scale=None, group=None, ops: Optional[BlockOps] = None):
    """Sequence-parallel attention.  q, k, v: this rank's LOCAL rows, (bh, n_local, d) or (B, H, n_local, d), head dim
    a multiple of 8 up to 128; for ``causal=True`` they must be laid out zig-zag (``zigzag_split``).  Returns (o, lse) for the local
    rows; differentiable."""
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    _check_equal_across_ranks(q, group)
    lead = q.shape[:-2]
    qb, kb, vb = (t.reshape(-1, t.shape[-2], t.shape[-1]) for t in (q, k, v))
    o, lse = _RingAttnFn.apply(qb, kb, vb, causal, softmax_scale, group, ops)
    return o.reshape(*lead, *o.shape[-2:]), lse.reshape(*lead, lse.shape[-1])
