"""Batch*head sharding (BASELINE config C4).

Every (batch, head) slice is an independent attention problem (the reference loops over them serially:
``csrc/fa2/fa2_fwd.cu:56``), so P ranks simply own contiguous ranges of the merged batch*head axis.  There is NO
data-path collective; ranks only meet at the timing barrier of the benchmark."""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the ``total`` slices owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def sharded_attention(attention_fn, q, k, v, rank: int, world: int, **kwargs):
    """Run ``attention_fn`` (e.g. ``fa2_attention``) on this rank's slice range of global (BH, N, d) tensors.

    Returns ((begin, end), (o_local, lse_local)).  In production each rank only ever materialises its own range; this
    helper exists so tests can check that the ranges tile the problem exactly."""
    begin, end = shard_range(q.shape[0], rank, world)
    return (begin, end), attention_fn(q[begin:end], k[begin:end], v[begin:end], **kwargs)
