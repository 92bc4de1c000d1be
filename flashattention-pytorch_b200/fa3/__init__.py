from .op import fa3_attention  # noqa: F401
from .spec import FA3Spec, pick_fa3_spec  # noqa: F401
from .sparse import fa3_block_sparse_attention  # noqa: F401
