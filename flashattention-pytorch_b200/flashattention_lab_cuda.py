"""``flashattention_lab_cuda`` — the module name the reference's CUDA wrappers import
(reference ``src/fa2/cuda/impl.py:6-16``), re-implemented as a thin ctypes shim over the hand-written sm_100a
library ``libfa_sm100.so`` (C ABI in ``include/fa_sm100.h``).

It exports exactly the six functions of the reference's pybind module
(reference ``csrc/common/torch.extension.cpp:73-83``) with the same positional arguments and return tuples:

    fa1_forward / forward / fa3_forward      (q, k, v, causal, softmax_scale, br, bc[, stages, fp8]) -> (o, lse)
    fa1_backward / backward / fa3_backward   (q, k, v, o, do, lse, causal, softmax_scale, br, bc[, stages, fp8])
                                             -> (dq, dk, dv)

Deliberate deviations from the reference (see DESIGN.md "Deviations"):
  * ``forward`` (FA2) returns the correctly normalised O; the reference divides by the row sum twice
    (reference ``csrc/fa2/fa2_fwd.cu:92-93,99``).
  * the backward uses the causal block rule of the Python twin (reference ``src/fa1/torch/impl.py:89``), not the
    inverted test in ``csrc/fa1/fa1_bwd.cu:80``.
  * ``br``/``bc``/``stages`` are accepted and ignored: the kernel picks its own tcgen05 tile shapes and the result
    does not depend on them beyond rounding.
  * fp32 inputs run in fp32 arithmetic on the CUDA cores (the reference's contract for fp32, tolerance 1e-4).
  * ``fp8=True`` (FA3) runs a real e4m3 tensor-core forward (head dim 128); the reference only emulates fp8, and that
    emulation is numerically broken, so parity is against this repo's own quantise -> dequantise oracle.
  * head dims that are a multiple of 8 (<= 128) run natively; others are zero-padded to the next multiple of 8.

There is NO fallback: if the shared library is missing or the device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
_LIB_NAME = "libfa_sm100.so"
_lib = None


class _Shape(ctypes.Structure):
    """Mirror of ``fa_sm100_shape`` (include/fa_sm100.h)."""

    _fields_ = [
        ("bh", ctypes.c_int64),
        ("n_q", ctypes.c_int64),
        ("n_kv", ctypes.c_int64),
        ("d", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("causal", ctypes.c_int32),
        ("softmax_scale", ctypes.c_float),
        ("q_row0", ctypes.c_int64),
        ("kv_col0", ctypes.c_int64),
        ("q_bh_stride", ctypes.c_int64),
        ("kv_bh_stride", ctypes.c_int64),
        ("lse_bh_stride", ctypes.c_int64),
    ]


class _Ext(ctypes.Structure):
    """Mirror of ``fa_sm100_ext`` (include/fa_sm100.h): block-sparse tile mask + dropout."""

    _fields_ = [
        ("block_mask", ctypes.c_void_p),
        ("mask_bh_stride", ctypes.c_int64),
        ("dropout_p", ctypes.c_float),
        ("seed", ctypes.c_uint64),
        ("offset", ctypes.c_uint64),
    ]


# every symbol include/fa_sm100.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_SP = ctypes.POINTER(_Shape)
_EP = ctypes.POINTER(_Ext)
ABI = {
    "fa_sm100_version": (ctypes.c_int, []),
    "fa_sm100_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "fa_sm100_dq_accum_bytes": (ctypes.c_size_t, [_SP]),
    "fa_sm100_fwd": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_fwd_ex": (ctypes.c_int, [_SP, _EP, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_bwd_ex": (ctypes.c_int, [_SP, _EP, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_rowstats_bytes": (ctypes.c_size_t, [_SP]),
    "fa_sm100_bwd_prepare": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_bwd": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_bwd_accum": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_int32, _P]),
    "fa_sm100_fwd_f32": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_bwd_f32": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_fp8_quantize": (ctypes.c_int, [_P, _P, _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, _P]),
    "fa_sm100_fwd_fp8": (ctypes.c_int, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fa_sm100_dq_finish": (ctypes.c_int, [_SP, _P, _P, _P]),
    "fa_sm100_cast_scaled": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_float, ctypes.c_int32, _P]),
}


def library_path() -> Path:
    override = os.environ.get("FA_SM100_LIB")
    return Path(override) if override else _HERE / _LIB_NAME


def load_library():
    """dlopen libfa_sm100.so and bind the C ABI.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU/Triton fallback on this path."
        )
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.fa_sm100_version() < 100:
        raise ImportError("libfa_sm100.so is older than this shim")
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().fa_sm100_strerror(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


_DTYPES = {torch.float16: 0, torch.bfloat16: 1}
_F32 = 2  # FA_SM100_DTYPE_F32: only the *_f32 entry points take it


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise NotImplementedError(
            f"flashattention_lab_cuda (sm_100a): dtype {t.dtype} is not supported; the tcgen05 path takes fp16/bf16 "
            "(fp32 tensors go through fwd_f32_raw / bwd_f32_raw, which the six public functions pick automatically)"
        ) from None


MAX_HEAD_DIM = 128          # fp32, FP8, ring-accumulator and block-sparse / dropout forms
MAX_HEAD_DIM_FORWARD = 256  # the plain 16-bit forward and backward have dedicated 129..256 kernels


def _padded_head_dim(d: int, limit: int = MAX_HEAD_DIM) -> int:
    """Head dims that are a multiple of 8 go to the kernels as they are (the TMA tensor maps carry the true d: columns
    up to the kernel variant's 64 / 128 / 256 are zero-filled on load and clipped on store, no copies).  Other sizes —
    rows would not be 16-byte aligned — are zero-padded to the next multiple of 8 by the shim."""
    if d > limit:
        raise NotImplementedError(f"flashattention_lab_cuda (sm_100a): head dim {d} > {limit} is not supported by this path")
    if os.environ.get("FA_SM100_PAD_HEAD_DIM") == "1":  # debugging aid: the round-1 behaviour (pad to 64 / 128)
        return 64 if d <= 64 else 128
    return (d + 7) // 8 * 8


def _pad_d(x: torch.Tensor, dp: int) -> torch.Tensor:
    d = x.shape[-1]
    if d == dp:
        return x.contiguous()
    return torch.nn.functional.pad(x, (0, dp - d)).contiguous()


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _validate_qkv(q, k, v):
    # reference csrc/fa2/fa2_fwd.cu:40-45 (TORCH_CHECK -> RuntimeError)
    if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
        raise RuntimeError("q, k, v must be 3-D (batch*heads, seqlen, head_dim)")
    if q.shape != k.shape or q.shape != v.shape:
        raise RuntimeError("q, k, v must have identical shapes")
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("Inputs must be CUDA tensors")
    if not (q.dtype == k.dtype == v.dtype):
        raise RuntimeError("q, k, v must share one dtype")


def make_shape(bh, n_q, n_kv, d, dtype_code, causal, softmax_scale, q_row0=0, kv_col0=0, q_bh_stride=0,
               kv_bh_stride=0, lse_bh_stride=0) -> _Shape:
    return _Shape(int(bh), int(n_q), int(n_kv), int(d), int(dtype_code), 1 if causal else 0, float(softmax_scale),
                  int(q_row0), int(kv_col0), int(q_bh_stride), int(kv_bh_stride), int(lse_bh_stride))


# ------------------------------------------------------------------------------------------------------------------
# raw entry points (head dim a multiple of 8, <= 128) — also used by the sharding and ring-attention drivers.
# Tensors may be views of a larger (bh, n, d) tensor along n: rows must be dense, the slice stride is free.
# ------------------------------------------------------------------------------------------------------------------
def _slice_stride(t: torch.Tensor) -> int:
    if t.dim() == 3:
        if t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            raise RuntimeError("rows must be dense (stride(1) == head_dim, stride(2) == 1)")
    elif t.dim() == 2:
        if t.stride(1) != 1:
            raise RuntimeError("lse rows must be dense")
    dense = t.shape[1] * (t.shape[2] if t.dim() == 3 else 1)
    if t.shape[0] > 1 and t.stride(0) < dense:  # expanded / overlapping slices: 0 would mean "dense" to the C ABI
        raise RuntimeError("slices must not overlap (stride(0) >= rows * head_dim); call .contiguous() first")
    return t.stride(0) if t.shape[0] > 1 else dense


def _same_stride(ref: torch.Tensor, *others: torch.Tensor) -> int:
    st = _slice_stride(ref)
    for o in others:
        if o.shape[0] > 1 and _slice_stride(o) != st:
            raise RuntimeError("tensors that share a geometry must share their slice stride")
    return st


def _empty_like_strided(t: torch.Tensor, dtype=None) -> torch.Tensor:
    return torch.empty_strided(t.shape, t.stride(), dtype=dtype or t.dtype, device=t.device)


def fwd_raw(q, k, v, causal, softmax_scale, *, q_row0=0, kv_col0=0, out=None, lse=None, merge=False):
    """q: (bh, n_q, d), k/v: (bh, n_kv, d), d % 8 == 0, d <= 256.  Returns (o, lse).

    ``merge=True`` folds the new partial into the given ``out``/``lse`` by log-sum-exp (ring attention)."""
    lib = load_library()
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    if out is None:
        if merge:
            raise ValueError("merge=True needs out/lse from the previous step")
        out = _empty_like_strided(q)
        lse = torch.empty((bh, n_q), device=q.device, dtype=torch.float32)
    elif lse is None:
        raise ValueError("out= needs lse= as well")
    shape = make_shape(bh, n_q, n_kv, d, _dtype_code(q), causal, softmax_scale, q_row0, kv_col0,
                       _same_stride(q, out), _same_stride(k, v), _slice_stride(lse))
    prev_o = out.data_ptr() if merge else None
    prev_lse = lse.data_ptr() if merge else None
    with torch.cuda.device(q.device):
        rc = lib.fa_sm100_fwd(ctypes.byref(shape), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                              lse.data_ptr(), prev_o, prev_lse, _stream_ptr(q))
    _check(rc, "fa_sm100_fwd")
    return out, lse


def bwd_prepare_raw(o, do, lse, zero=None):
    """Pre-pass of the backward: packs (-lse * log2e, -delta = -rowsum(dO o O)) per 128-row query tile.
    ``zero``: an fp32 dQ accumulator with o's shape and slice stride to zero-fill in the same launch."""
    lib = load_library()
    bh, n_q, d = o.shape
    stride = _same_stride(o, do) if zero is None else _same_stride(o, do, zero)
    if zero is not None and (zero.dtype != torch.float32 or zero.shape != o.shape):
        raise ValueError("zero= must be an fp32 tensor with o's shape")
    shape = make_shape(bh, n_q, n_q, d, _dtype_code(o), False, 1.0, 0, 0, stride, 0, _slice_stride(lse))
    rowstats = torch.empty(lib.fa_sm100_rowstats_bytes(ctypes.byref(shape)) // 4, device=o.device,
                           dtype=torch.float32)
    with torch.cuda.device(o.device):
        _check(lib.fa_sm100_bwd_prepare(ctypes.byref(shape), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                        rowstats.data_ptr(), None if zero is None else zero.data_ptr(),
                                        _stream_ptr(o)), "fa_sm100_bwd_prepare")
    return rowstats


def bwd_raw(q, k, v, o, do, lse, causal, softmax_scale, *, q_row0=0, kv_col0=0, rowstats=None, dq_accum=None,
            dk_accum=None, dv_accum=None, accum_overwrite=False):
    """Backward (d % 8 == 0, d <= 256; the ring forms with accumulators: d <= 128).

    Plain call: returns (dq, dk, dv) in the input dtype.
    Ring call (``dq_accum`` fp32 with q's shape AND strides given, ``rowstats`` from ``bwd_prepare_raw``): dQ partials
    are added into ``dq_accum`` (finish with ``dq_finish_raw`` after the last step); returns (None, dk, dv) of THIS
    K/V block (``o``/``lse`` may then be None).  With ``dk_accum``/``dv_accum`` (fp32, k's shape) the dK/dV partials go to
    them in fp32 instead — reduce-added, or stored with ``accum_overwrite=True`` — and (None, None, None) is returned."""
    lib = load_library()
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    ring = dq_accum is not None
    if not ring:
        if bh > 1 and _slice_stride(q) != n_q * d:
            q, do, o = q.contiguous(), do.contiguous(), o.contiguous()
        dq_accum = torch.empty(q.shape, device=q.device, dtype=torch.float32)
        if rowstats is None:
            rowstats = bwd_prepare_raw(o, do, lse, zero=dq_accum)  # one launch: statistics + zero-fill
        else:
            dq_accum.zero_()
    elif rowstats is None:
        rowstats = bwd_prepare_raw(o, do, lse)
    if (dk_accum is None) != (dv_accum is None):
        raise ValueError("dk_accum and dv_accum go together")
    if dk_accum is not None:
        if not ring:
            raise ValueError("dk_accum/dv_accum need dq_accum (ring call)")
        for t in (dk_accum, dv_accum):
            if t.dtype != torch.float32 or t.shape != k.shape:
                raise ValueError("dk_accum/dv_accum must be fp32 with k's shape")
        shape = make_shape(bh, n_q, n_kv, d, _dtype_code(q), causal, softmax_scale, q_row0, kv_col0,
                           _same_stride(q, do, dq_accum), _same_stride(k, v), 0)
        with torch.cuda.device(q.device):
            _check(lib.fa_sm100_bwd_accum(ctypes.byref(shape), q.data_ptr(), k.data_ptr(), v.data_ptr(), do.data_ptr(),
                                          rowstats.data_ptr(), dq_accum.data_ptr(), dk_accum.data_ptr(),
                                          dv_accum.data_ptr(), _same_stride(dk_accum, dv_accum),
                                          1 if accum_overwrite else 0, _stream_ptr(q)),
                   "fa_sm100_bwd_accum")
        return None, None, None
    dk = _empty_like_strided(k)
    dv = _empty_like_strided(k)
    shape = make_shape(bh, n_q, n_kv, d, _dtype_code(q), causal, softmax_scale, q_row0, kv_col0,
                       _same_stride(q, do, dq_accum), _same_stride(k, v, dk, dv), 0)
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_bwd(ctypes.byref(shape), q.data_ptr(), k.data_ptr(), v.data_ptr(), do.data_ptr(),
                                rowstats.data_ptr(), dq_accum.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                _stream_ptr(q)), "fa_sm100_bwd")
    if ring:
        return None, dk, dv
    return dq_finish_raw(dq_accum, q.dtype, softmax_scale), dk, dv


BLOCK = 128  # tile edge of the block-sparse mask (the reference's default block_size)


def _make_ext(block_mask, bh, n_q, n_kv, dropout_p, seed, offset):
    """(ctypes struct, tensors to keep alive).  ``block_mask``: (ceil(n_q/128), ceil(n_kv/128)) shared by all slices or
    (bh, ...) per slice, any integer / bool dtype, nonzero = compute the tile."""
    keep = None
    ptr, stride = None, 0
    if block_mask is not None:
        rows, cols = (n_q + BLOCK - 1) // BLOCK, (n_kv + BLOCK - 1) // BLOCK
        if block_mask.dim() == 2:
            want = (rows, cols)
        elif block_mask.dim() == 3:
            want, stride = (bh, rows, cols), rows * cols
        else:
            raise RuntimeError("block_sparse_mask must be (q_blocks, k_blocks) or (batch*heads, q_blocks, k_blocks)")
        if tuple(block_mask.shape) != want:
            raise RuntimeError(f"block_sparse_mask has shape {tuple(block_mask.shape)}, expected {want} "
                               f"(block size {BLOCK})")
        keep = (block_mask != 0).to(torch.uint8).contiguous()
        ptr = keep.data_ptr()
    if not (0.0 <= float(dropout_p) < 1.0):
        raise ValueError("dropout_p must be in [0, 1)")
    return _Ext(ptr, stride, float(dropout_p), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1)), keep


def fwd_ex_raw(q, k, v, causal, softmax_scale, *, block_mask=None, dropout_p=0.0, seed=0, offset=0, q_row0=0,
               kv_col0=0):
    """Forward with a 128 x 128 block-sparse mask and / or dropout (C ABI ``fa_sm100_fwd_ex``)."""
    lib = load_library()
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    out = _empty_like_strided(q)
    lse = torch.empty((bh, n_q), device=q.device, dtype=torch.float32)
    ext, keep = _make_ext(block_mask, bh, n_q, n_kv, dropout_p, seed, offset)
    shape = make_shape(bh, n_q, n_kv, d, _dtype_code(q), causal, softmax_scale, q_row0, kv_col0,
                       _same_stride(q, out), _same_stride(k, v), _slice_stride(lse))
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_fwd_ex(ctypes.byref(shape), ctypes.byref(ext), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                   out.data_ptr(), lse.data_ptr(), _stream_ptr(q)), "fa_sm100_fwd_ex")
    if keep is not None:
        keep.record_stream(torch.cuda.current_stream(q.device))
    return out, lse


def bwd_ex_raw(q, k, v, o, do, lse, causal, softmax_scale, *, block_mask=None, dropout_p=0.0, seed=0, offset=0,
               q_row0=0, kv_col0=0):
    """Backward matching ``fwd_ex_raw`` (same mask, dropout_p, seed, offset).  Returns (dq, dk, dv)."""
    lib = load_library()
    bh, n_q, d = q.shape
    n_kv = k.shape[1]
    if bh > 1 and _slice_stride(q) != n_q * d:
        q, do, o = q.contiguous(), do.contiguous(), o.contiguous()
    dq_accum = torch.empty(q.shape, device=q.device, dtype=torch.float32)
    rowstats = bwd_prepare_raw(o, do, lse, zero=dq_accum)
    dk, dv = _empty_like_strided(k), _empty_like_strided(k)
    ext, keep = _make_ext(block_mask, bh, n_q, n_kv, dropout_p, seed, offset)
    shape = make_shape(bh, n_q, n_kv, d, _dtype_code(q), causal, softmax_scale, q_row0, kv_col0,
                       _same_stride(q, do, dq_accum), _same_stride(k, v, dk, dv), 0)
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_bwd_ex(ctypes.byref(shape), ctypes.byref(ext), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                   do.data_ptr(), rowstats.data_ptr(), dq_accum.data_ptr(), dk.data_ptr(),
                                   dv.data_ptr(), _stream_ptr(q)), "fa_sm100_bwd_ex")
    if keep is not None:
        keep.record_stream(torch.cuda.current_stream(q.device))
    return dq_finish_raw(dq_accum, q.dtype, softmax_scale), dk, dv


def fwd_f32_raw(q, k, v, causal, softmax_scale, *, q_row0=0, kv_col0=0):
    """fp32 tensors (d % 4 == 0, d <= 128): fp32 arithmetic on the CUDA cores, results in fp32."""
    lib = load_library()
    bh, n_q, d = q.shape
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    out = torch.empty_like(q)
    lse = torch.empty((bh, n_q), device=q.device, dtype=torch.float32)
    shape = make_shape(bh, n_q, k.shape[1], d, _F32, causal, softmax_scale, q_row0, kv_col0)
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_fwd_f32(ctypes.byref(shape), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                    lse.data_ptr(), _stream_ptr(q)), "fa_sm100_fwd_f32")
    return out, lse


def bwd_f32_raw(q, k, v, o, do, lse, causal, softmax_scale, *, q_row0=0, kv_col0=0):
    lib = load_library()
    bh, n_q, d = q.shape
    q, k, v, o, do = (t.contiguous() for t in (q, k, v, o, do))
    lse = lse.to(torch.float32).contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty((bh, n_q), device=q.device, dtype=torch.float32)
    shape = make_shape(bh, n_q, k.shape[1], d, _F32, causal, softmax_scale, q_row0, kv_col0)
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_bwd_f32(ctypes.byref(shape), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                    do.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                    dv.data_ptr(), _stream_ptr(q)), "fa_sm100_bwd_f32")
    return dq, dk, dv


FP8_SEED = 0  # the reference's incoherent-processing seed (src/fa3/torch/impl.py:123: seed=0)


def fp8_quantize_raw(x, hadamard, seed=FP8_SEED):
    """x: (bh, n, 128) fp16/bf16 -> (e4m3 bytes as a uint8 tensor (bh, n, 128), fp32 scales (bh, ceil(n/128)))."""
    lib = load_library()
    bh, n, d = x.shape
    if d != 128:
        raise NotImplementedError("the FP8 path supports head dim 128 only")
    x = x.contiguous()
    out = torch.empty((bh, n, d), device=x.device, dtype=torch.uint8)
    scales = torch.empty((bh, (n + 127) // 128), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _check(lib.fa_sm100_fp8_quantize(x.data_ptr(), out.data_ptr(), scales.data_ptr(), bh, n, d, 0, _dtype_code(x),
                                         1 if hadamard else 0, int(seed), _stream_ptr(x)), "fa_sm100_fp8_quantize")
    return out, scales


def fwd_fp8_raw(q, k, v, causal, softmax_scale, seed=FP8_SEED):
    """FP8 forward: quantise (Q, K with the Hadamard rotation; V plain) and run both products in e4m3 on the tensor
    cores.  q, k, v: (bh, n, 128) fp16/bf16.  Returns (o in q's dtype, lse fp32)."""
    lib = load_library()
    bh, n_q, d = q.shape
    q8, sq = fp8_quantize_raw(q, True, seed)
    k8, sk = fp8_quantize_raw(k, True, seed)
    v8, sv = fp8_quantize_raw(v, False)
    sv_ref = sv.amax(dim=1).contiguous()
    o = torch.empty_like(q, memory_format=torch.contiguous_format)
    lse = torch.empty((bh, n_q), device=q.device, dtype=torch.float32)
    shape = make_shape(bh, n_q, k.shape[1], d, _dtype_code(q), causal, softmax_scale)
    with torch.cuda.device(q.device):
        _check(lib.fa_sm100_fwd_fp8(ctypes.byref(shape), q8.data_ptr(), k8.data_ptr(), v8.data_ptr(), sq.data_ptr(),
                                    sk.data_ptr(), sv.data_ptr(), sv_ref.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                    _stream_ptr(q)), "fa_sm100_fwd_fp8")
    return o, lse


def dq_finish_raw(dq_accum, dtype, softmax_scale):
    lib = load_library()
    bh, n_q, d = dq_accum.shape
    dq = _empty_like_strided(dq_accum, dtype)
    shape = make_shape(bh, n_q, n_q, d, _DTYPES[dtype], False, softmax_scale, 0, 0, _same_stride(dq_accum, dq))
    with torch.cuda.device(dq_accum.device):
        _check(lib.fa_sm100_dq_finish(ctypes.byref(shape), dq_accum.data_ptr(), dq.data_ptr(),
                                      _stream_ptr(dq_accum)), "fa_sm100_dq_finish")
    return dq


def cast_scaled(acc: torch.Tensor, alpha: float, dtype: torch.dtype) -> torch.Tensor:
    lib = load_library()
    out = torch.empty(acc.shape, device=acc.device, dtype=dtype)
    with torch.cuda.device(acc.device):
        _check(lib.fa_sm100_cast_scaled(acc.data_ptr(), out.data_ptr(), acc.numel(), float(alpha), _DTYPES[dtype],
                                        _stream_ptr(acc)), "fa_sm100_cast_scaled")
    return out


# ------------------------------------------------------------------------------------------------------------------
# the six functions of the reference's pybind module
# ------------------------------------------------------------------------------------------------------------------
def _pad4(x):
    d = x.shape[-1]
    return x.contiguous() if d % 4 == 0 else torch.nn.functional.pad(x, (0, 4 - d % 4)).contiguous()


@torch.no_grad()  # reference csrc/fa2/fa2_fwd.cu:38 (NoGradGuard)
def _forward(q, k, v, causal, softmax_scale):
    _validate_qkv(q, k, v)
    d = q.shape[-1]
    if q.dtype == torch.float32:  # the reference's fp32 contract (csrc/fa1/fa1_fwd.cu:67): fp32 arithmetic end to end
        if d > MAX_HEAD_DIM:
            raise NotImplementedError(f"flashattention_lab_cuda (sm_100a): head dim {d} > {MAX_HEAD_DIM} is not supported")
        o, lse = fwd_f32_raw(_pad4(q), _pad4(k), _pad4(v), bool(causal), float(softmax_scale))
        return (o if d % 4 == 0 else o[..., :d].contiguous()), lse
    dp = _padded_head_dim(d, MAX_HEAD_DIM_FORWARD)
    o, lse = fwd_raw(_pad_d(q, dp), _pad_d(k, dp), _pad_d(v, dp), bool(causal), float(softmax_scale))
    if dp != d:
        o = o[..., :d].contiguous()
    return o, lse


@torch.no_grad()
def _backward(q, k, v, o, do, lse, causal, softmax_scale):
    _validate_qkv(q, k, v)
    if o.shape != q.shape or do.shape != q.shape:
        raise RuntimeError("o and do must have q's shape")
    if lse.shape != q.shape[:2]:
        raise RuntimeError("lse must be (batch*heads, seqlen)")
    d = q.shape[-1]
    if q.dtype == torch.float32:
        dq, dk, dv = bwd_f32_raw(_pad4(q), _pad4(k), _pad4(v), _pad4(o.float()), _pad4(do.float()), lse, bool(causal),
                                 float(softmax_scale))
        if d % 4:
            dq, dk, dv = (t[..., :d].contiguous() for t in (dq, dk, dv))
        return dq, dk, dv
    dp = _padded_head_dim(d, MAX_HEAD_DIM_FORWARD)
    dq, dk, dv = bwd_raw(_pad_d(q, dp), _pad_d(k, dp), _pad_d(v, dp), _pad_d(o, dp), _pad_d(do.to(q.dtype), dp),
                         lse.to(torch.float32).contiguous(), bool(causal), float(softmax_scale))
    if dp != d:
        dq, dk, dv = (t[..., :d].contiguous() for t in (dq, dk, dv))
    return dq, dk, dv


def fa1_forward(q, k, v, causal, softmax_scale, br, bc):
    return _forward(q, k, v, causal, softmax_scale)


def fa1_backward(q, k, v, o, do, lse, causal, softmax_scale, br, bc):
    return _backward(q, k, v, o, do, lse, causal, softmax_scale)


def forward(q, k, v, causal, softmax_scale, br, bc):
    return _forward(q, k, v, causal, softmax_scale)


def backward(q, k, v, o, do, lse, causal, softmax_scale, br, bc):
    return _backward(q, k, v, o, do, lse, causal, softmax_scale)


def fa3_forward(q, k, v, causal, softmax_scale, br, bc, stages, fp8):
    if fp8:  # real e4m3 path (the reference only emulates it, src/fa3/torch/impl.py:123-131); head dim 128, 16-bit inputs
        _validate_qkv(q, k, v)
        if q.shape[-1] != 128 or q.dtype not in _DTYPES:
            raise NotImplementedError("fp8=True needs fp16/bf16 inputs with head dim 128")
        with torch.no_grad():
            return fwd_fp8_raw(q, k, v, bool(causal), float(softmax_scale))
    return _forward(q, k, v, causal, softmax_scale)


def fa3_backward(q, k, v, o, do, lse, causal, softmax_scale, br, bc, stages, fp8):
    # fp8=True: the forward ran in e4m3; gradients are taken through the 16-bit kernels on the original q, k, v with the
    # forward's o / lse (straight-through, as the reference's fa3_backward does: csrc/fa3/fa3_bwd.cu ignores fp8)
    return _backward(q, k, v, o, do, lse, causal, softmax_scale)
