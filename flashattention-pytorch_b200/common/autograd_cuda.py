"""Shared machinery behind ``_FA1CudaFn`` / ``_FA2CudaFn`` / ``_FA3CudaFn``.

The reference keeps three near-identical ``torch.autograd.Function`` classes
(``src/fa{1,2,3}/cuda/impl.py:38-82``).  Their contract — which this module preserves — is:

    forward(ctx, q, k, v, causal, softmax_scale, br, bc[, stages, fp8]) -> (o, lse)
        raises RuntimeError("Inputs must be CUDA tensors") for CPU tensors; forces q/k/v contiguous;
        saves (q, k, v, o, lse) and the scalars on ctx.
    backward(ctx, do, dlse) -> (dq, dk, dv, None x (number of non-tensor forward args))
        dlse is ignored (as in the reference); do is forced contiguous.

The native module is looked up under the reference's names (``flashattention_lab_cuda`` first, then
``flashattention_lab._C``; reference ``src/fa2/cuda/impl.py:6-16``) — here that is the ctypes shim over
libfa_sm100.so.  A missing module is an ImportError; nothing is caught and there is no other backend.
"""
from __future__ import annotations

from importlib import import_module

_EXT = None
_EXT_NAMES = ("flashattention_lab_cuda", "flashattention_lab._C")


def load_ext():
    global _EXT
    if _EXT is None:
        errors = []
        for name in _EXT_NAMES:
            try:
                _EXT = import_module(name)
                break
            except ImportError as exc:  # keep looking, but remember why
                errors.append(f"{name}: {exc}")
        else:
            raise ImportError("CUDA extension module not found (" + "; ".join(errors) + ")")
    return _EXT


def require_cuda(*tensors):
    if not all(t.is_cuda for t in tensors):
        raise RuntimeError("Inputs must be CUDA tensors")


def run_forward(ctx, fwd_name, q, k, v, causal, softmax_scale, tile_args):
    """Common body of ``_FAnCudaFn.forward``; ``tile_args`` = (br, bc) or (br, bc, stages, fp8)."""
    ext = load_ext()
    require_cuda(q, k, v)
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o, lse = getattr(ext, fwd_name)(q, k, v, bool(causal), float(softmax_scale), *tile_args)
    ctx.save_for_backward(q, k, v, o, lse)
    ctx.causal = bool(causal)
    ctx.softmax_scale = float(softmax_scale)
    return o, lse


def run_backward(ctx, bwd_name, do, tile_args):
    ext = load_ext()
    q, k, v, o, lse = ctx.saved_tensors
    return getattr(ext, bwd_name)(q, k, v, o, do.contiguous(), lse, ctx.causal, ctx.softmax_scale, *tile_args)
