"""``fa2_attention`` — public entry point, signature as the reference's ``src/fa2/op.py:7``.

The dispatch is collapsed to the single sm_100a CUDA extension (north star): ``backend`` may be "auto" or "cuda";
"triton"/"torch" no longer exist on this path and raise ``ValueError`` like any unknown backend did in the reference
(``src/fa2/op.py:29``).  Unlike the reference's ``auto`` (``:14-19``) no exception from the CUDA path is swallowed."""
from .cuda.impl import fa2_cuda
from .spec import pick_fa2_spec


def fa2_attention(q, k, v, causal=False, softmax_scale=None, backend="auto"):
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** -0.5
    if backend not in ("auto", "cuda"):
        raise ValueError(backend)
    spec = pick_fa2_spec(q.shape[-1])
    return fa2_cuda(q, k, v, causal, softmax_scale, spec)
