"""Tile spec of the FA3 entry point: the record type and the values of the reference (``src/fa3/spec.py:3-13``).

The values travel through the extension ABI unchanged (``br``, ``bc``, ``stages``); the sm_100a kernel ignores them and uses
its own tcgen05 shapes (128-row query tiles, 128-row KV tiles), which does not change the result beyond rounding.
Two records exist, one per head-dim class, chosen at 64 like the reference does."""
from dataclasses import make_dataclass

FA3Spec = make_dataclass("FA3Spec", [("br", int), ("bc", int), ("num_warps", int), ("stages", int)], frozen=True)
FA3Spec.__module__ = __name__

_HEAD_DIM_UP_TO_64 = FA3Spec(128, 128, 8, 2)
_HEAD_DIM_ABOVE_64 = FA3Spec(64, 128, 8, 2)


def pick_fa3_spec(head_dim: int) -> FA3Spec:
    """Tile configuration for a head dimension (reference rule: square 128 tiles up to d = 64, 64 x 128 above)."""
    return _HEAD_DIM_UP_TO_64 if head_dim <= 64 else _HEAD_DIM_ABOVE_64
