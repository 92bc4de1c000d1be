from .op import fa1_attention  # noqa: F401
from .spec import FA1Spec, pick_fa1_spec  # noqa: F401
