"""Debug-only bindings of ``libfa_sm100_probes.so`` (C ABI in ``include/fa_sm100_probes.h``): descriptor bring-up
self-tests and hardware rate probes used by ``tools/`` and one GPU test.  Not imported by the product path."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
_P = ctypes.c_void_p
ABI = {
    "fa_sm100_probe_umma": (ctypes.c_int, [ctypes.c_int, ctypes.c_int32, _P, _P, _P, _P]),
    "fa_sm100_probe_reduce_rate": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    "fa_sm100_probe_ex2_rate": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P]),
    "fa_sm100_probe_mma_rate": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
}
_DTYPES = {torch.float16: 0, torch.bfloat16: 1}
_lib = None


def library_path() -> Path:
    override = os.environ.get("FA_SM100_PROBES_LIB")
    return Path(override) if override else _HERE / "libfa_sm100_probes.so"


def load_library():
    global _lib
    if _lib is None:
        path = library_path()
        if not path.exists():
            raise ImportError(f"{path} not found: build it with `python __graft_entry__.py`")
        lib = ctypes.CDLL(str(path))
        for name, (restype, argtypes) in ABI.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc})")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def probe_umma(mode: int, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    if int(mode) in (6, 7):  # e4m3 operands
        if a.dtype != torch.float8_e4m3fn or tuple(a.shape) != (128, 128) or tuple(b.shape) != (128, 128):
            raise ValueError("probe_umma modes 6/7: need float8_e4m3fn a, b of shape [128,128]")
        out = torch.empty((128, 128), device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            _check(lib.fa_sm100_probe_umma(int(mode), 0, a.data_ptr(), b.data_ptr(), out.data_ptr(), _stream(a)),
                   "fa_sm100_probe_umma")
        return out
    rows = 256 if int(mode) >= 4 else 128  # modes 4/5 drive a CTA pair: A and out have 256 rows
    if tuple(a.shape) != (rows, 128) or tuple(b.shape) != (128, 128) or not (a.is_contiguous() and b.is_contiguous()):
        raise ValueError(f"probe_umma mode {mode}: need contiguous a [{rows},128] and b [128,128]")
    out = torch.empty((rows, 128), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _check(lib.fa_sm100_probe_umma(int(mode), _DTYPES[a.dtype], a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                       _stream(a)), "fa_sm100_probe_umma")
    return out


def probe_mma_rate(pair: bool, a_from_tmem: bool, n: int, groups: int, ctas: int, device="cuda") -> None:
    """Launch the tensor-core issue-rate probe on the current stream (the caller times it with CUDA events)."""
    lib = load_library()
    dev = torch.device(device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib.fa_sm100_probe_mma_rate(int(bool(pair)), int(bool(a_from_tmem)), int(n), int(groups), int(ctas),
                                           stream), "fa_sm100_probe_mma_rate")


def probe_reduce_rate(acc: torch.Tensor, nkt: int, flags: int = 0) -> None:
    """acc: [slices, nqt*128, 128] fp32, contiguous; every element grows by nkt (launch on the current stream).
    flags: 1 rotated walk, 2 red.global.v4 from registers instead of TMA reduce, 4 one CTA per SM."""
    lib = load_library()
    if acc.dtype != torch.float32 or acc.dim() != 3 or acc.shape[2] != 128 or acc.shape[1] % 128 or not acc.is_contiguous():
        raise ValueError("probe_reduce_rate: need contiguous fp32 acc [slices, nqt*128, 128]")
    with torch.cuda.device(acc.device):
        _check(lib.fa_sm100_probe_reduce_rate(acc.data_ptr(), acc.shape[0], acc.shape[1] // 128, int(nkt),
                                              int(flags), _stream(acc)), "fa_sm100_probe_reduce_rate")


def probe_ex2_rate(mode: int, iters: int, ctas: int, device="cuda") -> None:
    """Launch the MUFU exp2 rate probe on the current stream (0: f32, 1: f16x2, 2: bf16x2); the caller times it."""
    lib = load_library()
    dev = torch.device(device)
    sink = torch.zeros(1, device=dev)
    with torch.cuda.device(dev):
        _check(lib.fa_sm100_probe_ex2_rate(int(mode), int(iters), int(ctas), sink.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream), "fa_sm100_probe_ex2_rate")

