"""Tile spec of the FA1 entry point — same dataclass and values as the reference (``src/fa1/spec.py``).

The values travel through the extension ABI unchanged (``br``, ``bc``); the sm_100a kernel ignores them and uses its
own tcgen05 shapes (128-row query tiles, 128-row KV tiles), which does not change the result beyond rounding."""
from dataclasses import dataclass


@dataclass(frozen=True)
class FA1Spec:
    br: int
    bc: int
    num_warps: int


def pick_fa1_spec(head_dim: int) -> FA1Spec:
    return FA1Spec(br=128, bc=128, num_warps=8) if head_dim <= 64 else FA1Spec(br=64, bc=128, num_warps=8)
