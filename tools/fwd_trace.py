"""Per-CTA lifetime breakdown of the forward kernel (needs the -DFA_FWD_TRACE variant).
usage (under gpurun): FA_SM100_LIB=tools/_variants/lib_ftrace.so python tools/fwd_trace.py [n] [causal] [bh]"""
import ctypes
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]


def main():
    import torch
    import flashattention_lab_cuda as ext

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    causal = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    bh = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    d = 128
    lib = ext.load_library()
    fn = lib.fa_sm100_debug_fwd_trace
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    torch.manual_seed(0)
    q, k, v = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    for _ in range(3):
        ext.fwd_raw(q, k, v, causal, d ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ext.fwd_raw(q, k, v, causal, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    n_ctas = bh * ((n + 255) // 256)
    buf = (ctypes.c_longlong * (10 * n_ctas))()
    assert fn(buf, n_ctas) == n_ctas
    rows = [[buf[c * 10 + x] for x in range(10)] for c in range(n_ctas)]
    ns_per_clk = sum(r[9] - r[1] for r in rows) / float(sum(r[8] - r[2] for r in rows))
    t_first = min(r[1] for r in rows)
    t_last = max(r[9] for r in rows)
    print(f"fwd trace n={n} causal={causal} bh={bh}: {n_ctas} CTAs, launch {e0.elapsed_time(e1) * 1000:.1f} us by events, "
          f"{(t_last - t_first) / 1000:.1f} us first entry to last exit, sm clock {1000 / ns_per_clk:.0f} MHz")
    names = ["prologue (entry -> block sync)", "first scores (sync -> S(0) ready)", "main loop (S(0) -> last softmax step)",
             "tail (last softmax -> last PV done)", "epilogue (normalise, stage, TMA store)", "exit (sync, TMEM dealloc)"]
    segs = [(2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8)]
    for name, (a, b) in zip(names, segs):
        xs = [(r[b] - r[a]) * ns_per_clk for r in rows]
        print(f"  {name:45s} mean {sum(xs) / len(xs):8.0f} ns   min {min(xs):8.0f}   max {max(xs):8.0f}")
    life = [(r[9] - r[1]) for r in rows]
    print(f"  {'CTA lifetime (wall clock)':45s} mean {sum(life) / len(life):8.0f} ns")
    by_sm = defaultdict(list)
    for r in rows:
        by_sm[r[0]].append((r[1], r[9]))
    gaps, busy, idle_tail = [], 0, []
    for sm, lst in by_sm.items():
        lst.sort()
        for (a0, a1), (b0, b1) in zip(lst, lst[1:]):
            gaps.append(b0 - a1)
        busy += sum(b - a for a, b in lst)
        idle_tail.append(t_last - lst[-1][1])
    print(f"  {'gap between consecutive CTAs on one SM':45s} mean {sum(gaps) / max(len(gaps), 1):8.0f} ns   "
          f"min {min(gaps):8.0f}   max {max(gaps):8.0f}   ({len(by_sm)} SMs used)")
    span = (t_last - t_first) * len(by_sm)
    print(f"  SM occupancy by resident CTAs: {100.0 * busy / span:.1f}% of (SMs x span); mean idle tail per SM "
          f"{sum(idle_tail) / len(idle_tail) / 1000:.1f} us; start skew {max(l[0][0] for l in by_sm.values()) - t_first} ns")
    steps = [(r[5] - r[4]) * ns_per_clk for r in rows]
    print(f"  main-loop share of lifetime: {100.0 * sum(steps) / sum(life):.1f}%")


if __name__ == "__main__":
    main()
