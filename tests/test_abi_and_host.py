"""CPU: the C-ABI library loads and exports every symbol include/fa_sm100.h declares; host-side logic of the
reference-facing wrappers (no compute calls — there is no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "fa_sm100.h").read_text()
PROBES_HEADER = (ROOT / "include" / "fa_sm100_probes.h").read_text()


def _declared_symbols(header=HEADER):
    return sorted(set(re.findall(r"\b(fa_sm100_[a-z0-9_]+)\s*\(", header)) - {"fa_sm100_shape", "fa_sm100_strerror"}
                  | ({"fa_sm100_strerror"} if header is HEADER else set()))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for needed in ("fa_sm100_fwd", "fa_sm100_bwd", "fa_sm100_bwd_accum", "fa_sm100_bwd_prepare", "fa_sm100_dq_finish",
                   "fa_sm100_fwd_ex", "fa_sm100_bwd_ex", "fa_sm100_fwd_f32", "fa_sm100_bwd_f32", "fa_sm100_fp8_quantize",
                   "fa_sm100_fwd_fp8", "fa_sm100_strerror"):
        assert needed in syms
    assert not any("probe" in s for s in syms), "probe kernels belong to the debug library, not the product ABI"


def test_library_exports_every_declared_symbol():
    import flashattention_lab_cuda as ext

    path = ext.library_path()
    assert path.exists(), f"{path} missing: run `python __graft_entry__.py`"
    lib = ctypes.CDLL(str(path))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libfa_sm100.so does not export {name}"
    assert set(ext.ABI) == set(_declared_symbols()), "shim ABI table and header disagree"


def test_debug_library_exports_every_declared_probe():
    import probes

    path = probes.library_path()
    assert path.exists(), f"{path} missing: run `python __graft_entry__.py`"
    lib = ctypes.CDLL(str(path))
    declared = _declared_symbols(PROBES_HEADER)
    assert declared and all("probe" in s for s in declared)
    for name in declared:
        assert hasattr(lib, name), f"libfa_sm100_probes.so does not export {name}"
    assert set(probes.ABI) == set(declared)


def test_library_metadata_calls_work_without_gpu():
    import flashattention_lab_cuda as ext

    lib = ext.load_library()
    assert lib.fa_sm100_version() >= 100
    assert lib.fa_sm100_strerror(0) == b"ok"
    assert b"head dim" in lib.fa_sm100_strerror(-2)
    shape = ext.make_shape(bh=3, n_q=200, n_kv=200, d=128, dtype_code=1, causal=True, softmax_scale=0.1)
    assert lib.fa_sm100_dq_accum_bytes(ctypes.byref(shape)) == 3 * 200 * 128 * 4
    assert lib.fa_sm100_rowstats_bytes(ctypes.byref(shape)) == 3 * 2 * 256 * 4


def test_argument_validation_happens_before_any_cuda_call():
    import flashattention_lab_cuda as ext

    lib = ext.load_library()
    good = dict(bh=1, n_q=8, n_kv=8, d=64, dtype_code=0, causal=False, softmax_scale=1.0)
    fake = ctypes.c_void_p(256)  # aligned non-null dummy; validation must fail before it is touched

    def fwd(**over):
        s = ext.make_shape(**{**good, **over})
        return lib.fa_sm100_fwd(ctypes.byref(s), fake, fake, fake, fake, fake, None, None, None)

    assert fwd(dtype_code=7) == -1
    assert fwd(d=96) == 0 or fwd(d=96) <= -6  # multiples of 8 up to 128 pass validation (then fail on the fake device)
    assert fwd(d=100) == -2 and fwd(d=264) == -2 and fwd(d=4) == -2
    assert fwd(d=136) <= -6 and fwd(d=256) <= -6  # the plain forward goes up to 256
    sb = ext.make_shape(**{**good, "d": 256})
    assert lib.fa_sm100_bwd(ctypes.byref(sb), fake, fake, fake, fake, fake, fake, fake, fake, None) <= -6  # valid: 129..256 kernel
    assert lib.fa_sm100_bwd_accum(ctypes.byref(sb), fake, fake, fake, fake, fake, fake, fake, fake, 0, 1, None) == -2
    assert fwd(n_q=0) == -3
    assert fwd(softmax_scale=0.0) == -5
    s = ext.make_shape(**good)
    assert lib.fa_sm100_fwd(ctypes.byref(s), None, fake, fake, fake, fake, None, None, None) == -4
    assert lib.fa_sm100_fwd(ctypes.byref(s), ctypes.c_void_p(260), fake, fake, fake, fake, None, None, None) == -4
    # extras: dropout_p outside [0, 1), offsets that are not multiples of 4 with dropout on
    bad = ext._Ext(None, 0, 1.0, 0, 0)
    assert lib.fa_sm100_fwd_ex(ctypes.byref(s), ctypes.byref(bad), fake, fake, fake, fake, fake, None) == -9
    s2 = ext.make_shape(**{**good, "q_row0": 2})
    drop = ext._Ext(None, 0, 0.5, 1, 0)
    assert lib.fa_sm100_fwd_ex(ctypes.byref(s2), ctypes.byref(drop), fake, fake, fake, fake, fake, None) == -9
    assert b"extras" in lib.fa_sm100_strerror(-9)


def test_fp32_and_fp8_entries_validate_before_any_cuda_call():
    import flashattention_lab_cuda as ext

    lib = ext.load_library()
    fake = ctypes.c_void_p(256)
    f32 = dict(bh=1, n_q=8, n_kv=8, d=64, dtype_code=2, causal=False, softmax_scale=1.0)

    def fwd32(**over):
        s = ext.make_shape(**{**f32, **over})
        return lib.fa_sm100_fwd_f32(ctypes.byref(s), fake, fake, fake, fake, fake, None)

    assert fwd32(dtype_code=1) == -1          # the fp32 entries take fp32 only
    assert fwd32(d=130) == -2 and fwd32(d=6) == -2 and fwd32(d=256) == -2
    assert fwd32(n_kv=0) == -3
    assert fwd32() <= -6                      # valid arguments: fails only on the missing device
    s = ext.make_shape(**f32)
    assert lib.fa_sm100_fwd_f32(ctypes.byref(s), None, fake, fake, fake, fake, None) == -4
    assert lib.fa_sm100_bwd_f32(ctypes.byref(s), fake, fake, fake, fake, fake, fake, None, fake, fake, fake, None) == -4

    q = lambda **kw: lib.fa_sm100_fp8_quantize(fake, fake, fake, kw.get("bh", 2), kw.get("n", 256), kw.get("d", 128),
                                               kw.get("stride", 0), kw.get("dtype", 1), 1, 0, None)
    assert q(dtype=2) == -1 and q(d=64) == -2 and q(n=0) == -3 and q(stride=100) == -3
    assert q() <= -6
    s8 = ext.make_shape(bh=1, n_q=256, n_kv=256, d=64, dtype_code=1, causal=True, softmax_scale=0.1)
    assert lib.fa_sm100_fwd_fp8(ctypes.byref(s8), fake, fake, fake, fake, fake, fake, fake, fake, fake, None) == -2
    with pytest.raises(NotImplementedError):
        ext.fp8_quantize_raw(torch.zeros(1, 8, 64, dtype=torch.bfloat16), True)


def test_block_mask_argument_checks():
    import flashattention_lab_cuda as ext

    ok, keep = ext._make_ext(torch.ones(2, 3, dtype=torch.bool), 4, 200, 300, 0.25, 7, 4)
    assert keep.dtype == torch.uint8 and ok.mask_bh_stride == 0 and ok.seed == 7 and ok.offset == 4
    per_slice, _ = ext._make_ext(torch.ones(4, 2, 3), 4, 200, 300, 0.0, 0, 0)
    assert per_slice.mask_bh_stride == 6
    with pytest.raises(RuntimeError):
        ext._make_ext(torch.ones(3, 3), 4, 200, 300, 0.0, 0, 0)          # wrong tile grid
    with pytest.raises(RuntimeError):
        ext._make_ext(torch.ones(2, 2, 3), 4, 200, 300, 0.0, 0, 0)       # per-slice mask with the wrong slice count
    with pytest.raises(RuntimeError):
        ext._make_ext(torch.ones(6), 4, 200, 300, 0.0, 0, 0)
    for p in (-0.1, 1.0):
        with pytest.raises(ValueError):
            ext._make_ext(None, 4, 200, 300, p, 0, 0)


def test_shape_struct_layout_matches_header():
    import flashattention_lab_cuda as ext

    # 3*8 + 3*4 + 4 + 5*8 = 80 bytes with natural alignment
    assert ctypes.sizeof(ext._Shape) == 80
    def struct_fields(name):
        body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", HEADER, flags=re.S).group(1)
        return re.findall(r"^\s+(?:const\s+)?(?:u?int64_t|int32_t|float|uint8_t\*)\s+(\w+);", body, flags=re.M)

    assert [f[0] for f in ext._Shape._fields_] == struct_fields("fa_sm100_shape")
    # fa_sm100_ext: pointer, int64, float (+4 pad), uint64, uint64 = 40 bytes
    assert [f[0] for f in ext._Ext._fields_] == struct_fields("fa_sm100_ext")
    assert ctypes.sizeof(ext._Ext) == 40


@pytest.mark.parametrize("n", [1, 2, 3])
def test_specs_match_reference_values(n):
    mod = __import__(f"fa{n}.spec", fromlist=["x"])
    pick = getattr(mod, f"pick_fa{n}_spec")
    small, big = pick(64), pick(128)
    assert (small.br, small.bc, small.num_warps) == (128, 128, 8)  # reference src/faN/spec.py
    assert (big.br, big.bc, big.num_warps) == (64, 128, 8)
    if n == 3:
        assert small.stages == 2 and big.stages == 2
    with pytest.raises(Exception):
        small.br = 1  # frozen dataclass


def test_merge_split_roundtrip():
    from common.utils import merge_bh, split_bh, split_bh_lse

    x = torch.arange(2 * 3 * 4 * 5.0).reshape(2, 3, 4, 5)
    m, shp = merge_bh(x)
    assert m.shape == (6, 4, 5) and shp == (2, 3)
    assert torch.equal(split_bh(m, shp), x)
    assert split_bh_lse(torch.zeros(6, 4), shp).shape == (2, 3, 4)
    y, none = merge_bh(m)  # 3-D input: (tensor, None) — the reference's FA1/FA2 wrappers get this wrong (D6)
    assert none is None and y is m


@pytest.mark.parametrize("n", [1, 2, 3])
def test_cpu_tensors_are_rejected_like_the_reference(n):
    mod = __import__(f"fa{n}.op", fromlist=["x"])
    attn = getattr(mod, f"fa{n}_attention")
    q = torch.randn(1, 2, 8, 32)
    with pytest.raises(RuntimeError, match="Inputs must be CUDA tensors"):  # reference src/fa2/cuda/impl.py:43-44
        attn(q, q, q, backend="cuda")
    with pytest.raises(RuntimeError, match="Inputs must be CUDA tensors"):
        attn(q, q, q)  # "auto" no longer falls back to torch/triton
    with pytest.raises(ValueError):
        attn(q, q, q, backend="nope")  # reference src/fa2/op.py:29
    for gone in ("triton", "torch"):
        with pytest.raises(ValueError):
            attn(q, q, q, backend=gone)


def test_autograd_function_signatures():
    import inspect

    from fa1.cuda.impl import _FA1CudaFn
    from fa2.cuda.impl import _FA2CudaFn
    from fa3.cuda.impl import _FA3CudaFn

    base = ["ctx", "q", "k", "v", "causal", "softmax_scale", "br", "bc"]
    assert list(inspect.signature(_FA1CudaFn.forward).parameters) == base
    assert list(inspect.signature(_FA2CudaFn.forward).parameters) == base
    assert list(inspect.signature(_FA3CudaFn.forward).parameters) == base + ["stages", "fp8"]
    for fn in (_FA1CudaFn, _FA2CudaFn, _FA3CudaFn):
        assert list(inspect.signature(fn.backward).parameters) == ["ctx", "do", "dlse"]


def test_extension_module_exports_reference_names():
    import flashattention_lab_cuda as ext

    for name in ("fa1_forward", "fa1_backward", "forward", "backward", "fa3_forward", "fa3_backward"):
        assert callable(getattr(ext, name))  # reference csrc/common/torch.extension.cpp:73-83
    x = torch.randn(2, 16, 128)
    with pytest.raises(RuntimeError, match="CUDA"):  # fp8=True is a real path now: CPU tensors fail like everywhere else
        ext.fa3_forward(x, x, x, False, 1.0, 128, 128, 2, True)


def test_product_path_never_imports_the_oracle():
    pkg = ROOT / "flashattention-pytorch_b200"
    for py in pkg.rglob("*.py"):
        text = py.read_text()
        assert "oracle" not in text.replace("oracle/", "").replace("the oracle", "") or "import oracle" not in text, py
        assert "import oracle" not in text and "from oracle" not in text, f"{py} imports the oracle"
