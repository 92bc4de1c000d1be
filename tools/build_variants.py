"""Build libfa_sm100 variants with different -D flags into gpurun_out-independent dir tools/_variants/ (git-ignored *.so)."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g
out = ROOT / "tools" / "_variants"; out.mkdir(exist_ok=True)
variants = dict(a.split("=", 1) for a in sys.argv[1:])  # name=-DX=1,-DY=2
srcs = sorted(g.CSRC.glob("*.cu"))
procs = []
for name, flags in variants.items():
    fl = [f for f in flags.split(",") if f]
    objs = []
    for s in srcs:
        o = out / f"{name}_{s.stem}.o"; objs.append(o)
        procs.append(subprocess.Popen([g._nvcc(), *g.NVCC_FLAGS, *fl, "-c", str(s), "-o", str(o)]))
    variants[name] = objs
for p in procs:
    assert p.wait() == 0
for name, objs in variants.items():
    subprocess.run([g._nvcc(), *g.NVCC_FLAGS, "-shared", "-o", str(out / f"lib_{name}.so"), *map(str, objs)], check=True)
    for o in objs: o.unlink()
    print("built", name)
