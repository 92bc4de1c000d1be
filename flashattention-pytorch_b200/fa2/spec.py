"""Tile spec of the FA2 entry point: the record type and the values of the reference (``src/fa2/spec.py:3-12``).

The values travel through the extension ABI unchanged (``br``, ``bc``); the sm_100a kernel ignores them and uses
its own tcgen05 shapes (128-row query tiles, 128-row KV tiles), which does not change the result beyond rounding.
Two records exist, one per head-dim class, chosen at 64 like the reference does."""
from dataclasses import make_dataclass

FA2Spec = make_dataclass("FA2Spec", [("br", int), ("bc", int), ("num_warps", int)], frozen=True)
FA2Spec.__module__ = __name__

_HEAD_DIM_UP_TO_64 = FA2Spec(128, 128, 8)
_HEAD_DIM_ABOVE_64 = FA2Spec(64, 128, 8)


def pick_fa2_spec(head_dim: int) -> FA2Spec:
    """Tile configuration for a head dimension (reference rule: square 128 tiles up to d = 64, 64 x 128 above)."""
    return _HEAD_DIM_UP_TO_64 if head_dim <= 64 else _HEAD_DIM_ABOVE_64
