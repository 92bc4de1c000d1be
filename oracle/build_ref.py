"""Compile the REFERENCE's own native extension for CPU into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's csrc/fa*/*.cu contain no device code — they are host-side ATen loops — so g++ can
compile them as C++ straight from where they lie under /root/reference (nothing is copied into this repo).  The result
is a pybind11 module `flashattention_lab_cuda_ref` exporting the reference's six functions, used by
tests/test_oracle_golden.py to pin the oracle and by `bench.py --impl reference` as the reference arm.

Known defects of what this builds (documented, not fixed — it is the unmodified reference): FA2 `forward` divides
by the row sum twice (D2); every `*_backward` has the causal block skip inverted, so causal multi-tile gradients are
wrong (D4).  `csrc/common/bindings.cpp` is NOT compiled (second, conflicting PYBIND11_MODULE).

Only possible where /root/reference exists (the build container); the GPU box uses the prebuilt .so.
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

REF = Path("/root/reference/csrc")
OUT_DIR = Path(__file__).resolve().parent / "_ref"
NAME = "flashattention_lab_cuda_ref"
SOURCES = ["fa1/fa1_fwd.cu", "fa1/fa1_bwd.cu", "fa2/fa2_fwd.cu", "fa2/fa2_bwd.cu", "fa3/fa3_fwd.cu", "fa3/fa3_bwd.cu",
           "common/torch.extension.cpp"]


def main() -> int:
    if not REF.exists():
        print("[build_ref] /root/reference/csrc not present: nothing to do")
        return 0
    import torch
    from torch.utils import cpp_extension

    OUT_DIR.mkdir(exist_ok=True)
    target = OUT_DIR / f"{NAME}.so"
    srcs = [REF / s for s in SOURCES]
    if target.exists() and all(target.stat().st_mtime > s.stat().st_mtime for s in srcs):
        print(f"[build_ref] {target} up to date")
        return 0
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    lib_dir = Path(torch.__file__).parent / "lib"
    flags = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-w", f"-DTORCH_EXTENSION_NAME={NAME}",
             "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    objs, procs = [], []
    for s in srcs:
        obj = OUT_DIR / (s.stem + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", "-x", "c++", *flags, *inc, "-c", str(s), "-o", str(obj)]))
    if any(p.wait() != 0 for p in procs):
        print("[build_ref] compilation failed")
        return 1
    link = ["g++", "-shared", "-o", str(target), *map(str, objs), f"-L{lib_dir}", f"-Wl,-rpath,{lib_dir}",
            "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-fopenmp"]
    subprocess.run(link, check=True)
    for o in objs:
        o.unlink()
    print(f"[build_ref] built {target}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
