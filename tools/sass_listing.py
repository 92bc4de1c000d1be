"""SASS listings of the hot loops (north star: "committed SASS listings"): for the bf16 d=128 forward and backward
kernels of libfa_sm100.so this writes profiles/<tag>_sass_{fwd,bwd}_hotloop.txt with
  * an opcode census of the whole kernel (tensor-core, TMA, TMEM, MUFU, packed-fp32 and barrier instructions),
  * every basic block that issues UTCHMMA (the MMA issue loops), and
  * the blocks with the most MUFU.EX2 (the softmax / P recomputation inner loops),
straight from `cuobjdump -sass` (no GPU needed).   python tools/sass_listing.py [tag]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "flashattention-pytorch_b200" / "libfa_sm100.so"
KERNELS = {"fwd": "_ZN2fa13fa_fwd_kernelILi128ELb1ELb0EEE", "bwd": "_ZN2fa13fa_bwd_kernelILi128ELb1ELb0EEE"}  # D=128, bf16, dense
CENSUS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP",
          "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX", "F2FP", "SYNCS", "BAR.SYNC", "STS", "LDS", "RED", "ELECT",
          "R2UR", "WARPSYNC", "NANOSLEEP"]


def functions(sass: str):
    cur, buf = None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                yield cur, buf
            cur, buf = m.group(1), []
        elif cur is not None:
            buf.append(line)
    if cur:
        yield cur, buf


def blocks(lines):
    """Split a kernel's listing into basic blocks: a block starts at every branch target and after every branch."""
    addr = lambda ln: int(re.search(r"/\*([0-9a-f]{4,})\*/", ln).group(1), 16)  # noqa: E731
    targets = set()
    for ln in lines:
        m = re.search(r"\b(?:BRA|BRX|CALL|BSSY|BSYNC|WARPSYNC)\b.*?(0x[0-9a-f]+)\s*;", ln)
        if m and " BRA" in ln:
            targets.add(int(m.group(1), 16))
    out, cur = [], []
    for ln in lines:
        if addr(ln) in targets and cur:
            out.append(cur)
            cur = []
        cur.append(ln.rstrip())
        if re.search(r"\b(BRA|EXIT|RET)\b", ln):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def instr(line):
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    return m.group(1) if m else None


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = dict(functions(sass))
    for short, prefix in KERNELS.items():
        name = next(n for n in funcs if n.startswith(prefix))
        lines = [ln for ln in funcs[name] if instr(ln)]
        ins = [instr(ln) for ln in lines if instr(ln)]
        census = collections.Counter()
        for i in ins:
            op = i.split()[1] if i.startswith("@") else i.split()[0]
            for key in CENSUS:
                if op == key or op.startswith(key + "."):
                    census[key] += 1
        bl = blocks(lines)
        mma_blocks = [b for b in bl if any("UTCHMMA" in x for x in b)]
        ex2 = sorted(bl, key=lambda b: -sum("MUFU.EX2" in x for x in b))[:2]
        out = [f"# {name}", f"# source: cuobjdump -sass {LIB.relative_to(ROOT)} (sm_100a), {len(ins)} instructions",
               "#", "# opcode census (whole kernel):"]
        out += [f"#   {k:10s} {census[k]}" for k in CENSUS if census[k]]
        out += ["#", f"# ---- {len(mma_blocks)} basic blocks that issue UTCHMMA (tcgen05.mma): the MMA issue loops ----"]
        for b in mma_blocks:
            out += b + [""]
        out += ["# ---- the two blocks with the most MUFU.EX2 (exp2 of the softmax / P recomputation) ----"]
        for b in ex2:
            out += [f"# block with {sum('MUFU.EX2' in x for x in b)} MUFU.EX2, {sum('FFMA2' in x for x in b)} FFMA2, "
                    f"{sum('LDTM' in x for x in b)} LDTM, {sum('STTM' in x for x in b)} STTM"] + b + [""]
        path = ROOT / "profiles" / f"{tag}_sass_{short}_hotloop.txt"
        path.write_text("\n".join(out) + "\n")
        print(f"{path.relative_to(ROOT)}: {len(out)} lines; census {dict(census)}")


if __name__ == "__main__":
    main()
