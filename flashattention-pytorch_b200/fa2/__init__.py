from .op import fa2_attention  # noqa: F401
from .spec import FA2Spec, pick_fa2_spec  # noqa: F401
