"""Tile spec of the FA3 entry point — same dataclass and values as the reference (``src/fa3/spec.py``).

The values travel through the extension ABI unchanged (``br``, ``bc``, ``stages``); the sm_100a kernel ignores them and uses its
own tcgen05 shapes (128-row query tiles, 128-row KV tiles), which does not change the result beyond rounding."""
from dataclasses import dataclass


@dataclass(frozen=True)
class FA3Spec:
    br: int
    bc: int
    num_warps: int
    stages: int


def pick_fa3_spec(head_dim: int) -> FA3Spec:
    return FA3Spec(br=128, bc=128, num_warps=8, stages=2) if head_dim <= 64 else FA3Spec(br=64, bc=128, num_warps=8, stages=2)
