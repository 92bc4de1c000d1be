"""Multi-GPU drivers for the attention hot path: batch*head sharding (no collective) and ring attention."""
from .ring import (BlockOps, SymmMemRingDriver, TorchRingDriver, make_driver, contiguous_split, cuda_block_ops, ring_attention, ring_backward,  # noqa: F401
                   ring_forward, run_loopback, zigzag_chunk_ids, zigzag_merge, zigzag_split)
from .shard import shard_range, sharded_attention  # noqa: F401
