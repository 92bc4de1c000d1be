"""Per-phase timeline of one backward CTA (needs the -DFA_BWD_TRACE variant: tools/build_variants.py trace=-DFA_BWD_TRACE).

usage (under gpurun):  FA_SM100_LIB=tools/_variants/lib_trace.so python tools/bwd_trace.py [n] [causal] [bh]
Prints, per query-tile iteration, when each role passed each point (ns relative to the CTA's first event, assuming the
SM clock reported by nvidia-smi) and the steady-state period; writes gpurun_out/bwd_trace_<tag>.json.
"""
import ctypes
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]

EVENTS = {0: "X dV issue", 1: "X S(next) issue", 2: "Y dP issue", 3: "Y dK/dQ issue", 4: "C S ready", 5: "C P done",
          6: "C dP ready", 7: "C dS done", 8: "D dQ ready", 9: "D drained", 10: "D half0 read", 14: "D half1 issued",
          11: "D stage free", 12: "P Q load", 13: "P dO load"}
ORDER = [4, 5, 0, 1, 6, 7, 3, 8, 9, 2, 10, 14, 11, 12, 13]


def main():
    import torch
    import flashattention_lab_cuda as ext

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    causal = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    bh = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    d = 128
    lib = ext.load_library()
    fn = lib.fa_sm100_debug_bwd_trace
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    torch.manual_seed(0)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    for _ in range(2):
        ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5)
    torch.cuda.synchronize()
    nkt = n // 128
    # a middle slice; the heaviest kv tile when causal.  Linear block id of (slice, kv tile) under the library's grid
    # shape: x = slice inside its group, y = kv tile, z = group (csrc/fa_host.cuh sched_group_log2)
    lg = 0
    if causal:
        while (2 << lg) <= max(1, 296 // nkt) and (2 << lg) <= bh:
            lg += 1
    sl, jt = bh // 2, (0 if causal else nkt // 2)
    block = (sl & ((1 << lg) - 1)) + (1 << lg) * (jt + nkt * (sl >> lg))
    assert fn(block, None, 0) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    iters = 64
    buf = (ctypes.c_longlong * (16 * iters))()
    got = fn(-1, buf, 16 * iters)
    assert got == 16 * iters, got
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True).stdout.split()[0])
    ev = {e: [buf[e * iters + i] for i in range(iters)] for e in EVENTS}
    gt = [buf[15 * iters + i] for i in range(iters)]
    last = max(i for i in range(iters) if ev[4][i] > 0)
    if last > 8 and gt[last] > gt[0]:
        ns_per_clk = (gt[last] - gt[0]) / float(ev[4][last] - ev[4][0])
        mhz = 1000.0 / ns_per_clk  # measured under load from globaltimer
    else:
        ns_per_clk = 1000.0 / mhz
    t0 = min(x for xs in ev.values() for x in xs if x > 0)
    rel = {e: [(x - t0) * ns_per_clk if x > 0 else None for x in xs] for e, xs in ev.items()}
    n_it = sum(1 for x in ev[4] if x > 0)
    print(f"bwd trace n={n} causal={causal} bh={bh} block={block}: {n_it} iterations traced, whole launch "
          f"{e0.elapsed_time(e1):.3f} ms, sm clock {mhz:.0f} MHz")
    hdr = "it   " + " ".join(f"{EVENTS[e][:13]:>13s}" for e in ORDER)
    print(hdr)
    for it in range(min(n_it, 14)):
        print(f"{it:3d}  " + " ".join(f"{rel[e][it]:13.0f}" if rel[e][it] is not None else " " * 13 for e in ORDER))
    lo, hi = 4, min(n_it, iters) - 2
    if hi > lo:
        period = (rel[4][hi] - rel[4][lo]) / (hi - lo)
        print(f"steady-state period (S ready to S ready): {period:.0f} ns")

        def mean_gap(a, b, shift=0):
            xs = [rel[b][i + shift] - rel[a][i] for i in range(lo, hi - shift) if rel[a][i] is not None and rel[b][i + shift] is not None]
            return sum(xs) / len(xs) if xs else float("nan")

        gaps = {
            "S ready -> P done (P phase)": mean_gap(4, 5),
            "P done -> X dV issue (handoff)": mean_gap(5, 0),
            "X dV issue -> X S(next) issue": mean_gap(0, 1),
            "X S(next) issue -> S ready(next) (S MMA + handoff)": mean_gap(1, 4, 1),
            "P done -> dP ready (compute idle)": mean_gap(5, 6),
            "dP ready -> dS done (dS phase)": mean_gap(6, 7),
            "dS done -> Y dK/dQ issue (handoff)": mean_gap(7, 3),
            "Y dK/dQ issue -> D dQ ready (dK+dQ MMA + handoff)": mean_gap(3, 8),
            "D dQ ready -> drained": mean_gap(8, 9),
            "D drained -> Y dP issue(next) (handoff)": mean_gap(9, 2, 1),
            "Y dP issue -> dP ready (dP MMA + handoff)": mean_gap(2, 6),
            "D drained -> half0 read done": mean_gap(9, 10),
            "D half1 issued -> stage free (half1 read)": mean_gap(14, 11),
            "D stage free -> D dQ ready(next) (drain idle)": mean_gap(11, 8, 1),
            "dS done -> S ready(next) (compute idle)": mean_gap(7, 4, 1),
        }
        for name, val in gaps.items():
            print(f"  {name:55s} {val:8.0f} ns")
    else:
        gaps, period = {}, None
    # ---- per-CTA lifetime of every CTA of the same launch ----
    life_fn = getattr(lib, "fa_sm100_debug_bwd_life", None)
    if life_fn is not None:
        from collections import defaultdict
        life_fn.restype = ctypes.c_int
        life_fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
        gsz = 1 << lg
        n_ctas = min(16384, ((bh + gsz - 1) // gsz) * gsz * nkt)
        lb = (ctypes.c_longlong * (8 * n_ctas))()
        if life_fn(lb, n_ctas) == n_ctas:
            rows = [[lb[c * 8 + x] for x in range(8)] for c in range(n_ctas)]
            rows = [r for r in rows if r[7] > r[1] > 0]  # padding CTAs never stamp
            k = sum(r[7] - r[1] for r in rows) / float(sum(r[6] - r[2] for r in rows))
            t_first, t_last = min(r[1] for r in rows), max(r[7] for r in rows)
            print(f"lifetime of {len(rows)} CTAs: span {(t_last - t_first) / 1000:.1f} us")
            for name, (a, b) in (("prologue + first scores (entry -> S(0) ready)", (2, 3)),
                                 ("main loop (S(0) -> last dS done)", (3, 4)),
                                 ("tail (last dS -> last dK/dV product done)", (4, 5)),
                                 ("epilogue + exit (convert, stage, TMA store, dealloc)", (5, 6))):
                xs = [(r[b] - r[a]) * k for r in rows if r[b] > 0 and r[a] > 0]
                print(f"  {name:55s} mean {sum(xs) / len(xs):8.0f} ns   min {min(xs):8.0f}   max {max(xs):8.0f}")
            lifes = [r[7] - r[1] for r in rows]
            print(f"  {'CTA lifetime (wall clock)':55s} mean {sum(lifes) / len(lifes):8.0f} ns")
            by_sm = defaultdict(list)
            for r in rows:
                by_sm[r[0]].append((r[1], r[7]))
            gaps, idle_tail = [], []
            for lst in by_sm.values():
                lst.sort()
                gaps += [b0 - a1 for (a0, a1), (b0, b1) in zip(lst, lst[1:])]
                idle_tail.append(t_last - lst[-1][1])
            print(f"  {'gap between consecutive CTAs on one SM':55s} mean {sum(gaps) / max(len(gaps), 1):8.0f} ns")
            print(f"  mean idle tail per SM {sum(idle_tail) / len(idle_tail) / 1000:.1f} us; "
                  f"SM occupancy {100.0 * sum(lifes) / ((t_last - t_first) * len(by_sm)):.1f}%")
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    tag = f"n{n}_c{int(causal)}"
    (out / f"bwd_trace_{tag}.json").write_text(json.dumps({"n": n, "causal": causal, "bh": bh, "mhz": mhz,
                                                            "period_ns": period, "gaps_ns": gaps, "rel_ns": rel}))


if __name__ == "__main__":
    main()
