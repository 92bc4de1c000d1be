"""Per-phase timeline of one persistent backward CTA (needs the -DFA_BWD_TRACE variant:
python tools/build_variants.py trace=-DFA_BWD_TRACE).

usage (under gpurun):  FA_SM100_LIB=tools/_variants/lib_trace.so python tools/bwd_trace.py [n] [causal] [bh] [block]
Events are indexed by the CTA's running query-tile count, so item boundaries (the last tile of one K/V tile's walk, the
dK/dV write-out, the first tile of the next item) appear in line.  Prints the timeline around the first item boundaries,
the steady-state gaps between roles, the cost of an item boundary, and the wall-clock span of every CTA (load balance
of the static snake schedule).  Writes gpurun_out/bwd_trace_<tag>.json."""
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]

EVENTS = {0: "X dV issue", 1: "X S(next)", 2: "Y dP issue", 3: "Y dK/dQ", 4: "C S ready", 5: "C P done",
          6: "C dP ready", 7: "C dS done", 8: "D dQ ready", 9: "D drained", 10: "D half0 rd", 14: "D half1 is",
          11: "D stg free", 12: "P Q load", 13: "P dO load", 15: "C dKV done", 16: "C epi done", 17: "X KV ready"}
ORDER = [12, 13, 17, 4, 5, 0, 1, 2, 6, 7, 3, 8, 9, 11, 15, 16]
ITERS, NEV = 128, 20


def main():
    import torch
    import flashattention_lab_cuda as ext

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    causal = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    bh = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    block = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    d = 128
    lib = ext.load_library()
    fn = lib.fa_sm100_debug_bwd_trace
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    spans = lib.fa_sm100_debug_bwd_spans
    spans.restype, spans.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
    torch.manual_seed(0)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    acc = torch.empty(bh, n, d, device="cuda", dtype=torch.float32)
    stats = ext.bwd_prepare_raw(o, do, lse, zero=acc)
    run = lambda: ext.bwd_raw(q, k, v, None, do, None, causal, d ** -0.5, rowstats=stats, dq_accum=acc)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    assert fn(block, None, 0) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = (ctypes.c_longlong * (NEV * ITERS))()
    assert fn(-1, buf, NEV * ITERS) == NEV * ITERS
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    sp = (ctypes.c_longlong * (2 * 256))()
    n_ctas = min(256, sm)
    assert spans(sp, n_ctas) == n_ctas
    ev = {e: [buf[e * ITERS + i] for i in range(ITERS)] for e in EVENTS}
    stamped = [x for xs in ev.values() for x in xs if x > 0]
    t0 = min(stamped)
    starts = [sp[2 * c] for c in range(n_ctas) if sp[2 * c + 1] > sp[2 * c] > 0]
    ends = [sp[2 * c + 1] for c in range(n_ctas) if sp[2 * c + 1] > sp[2 * c] > 0]
    # SM clock from the traced CTA's wall-clock span vs its cycle span is not available per event; use the launch time:
    # cycles between the first and last stamp of the traced CTA over its wall-clock span
    span_ns = float(sp[2 * block + 1] - sp[2 * block]) if block < n_ctas else ms * 1e6
    ns_per_clk = span_ns / float(max(stamped) - t0) if max(stamped) > t0 else 0.5
    ns_per_clk = min(max(ns_per_clk, 0.45), 1.0)
    rel = {e: [(x - t0) * ns_per_clk if x > 0 else None for x in xs] for e, xs in ev.items()}
    n_it = sum(1 for x in ev[4] if x > 0)
    print(f"bwd trace n={n} causal={causal} bh={bh} block={block}: {n_it} query tiles traced, launch {ms:.3f} ms, "
          f"~{1000 / ns_per_clk:.0f} MHz assumed")
    boundaries = [i for i in range(n_it) if rel[16][i] is not None]
    print("item boundaries (last tile index of each item):", boundaries[:12])
    show = sorted(set(list(range(0, 6)) + [x for b in boundaries[:3] for x in range(max(b - 2, 0), b + 5)]))
    print("tile " + " ".join(f"{EVENTS[e][:10]:>10s}" for e in ORDER))
    for it in show:
        if it < n_it:
            print(f"{it:4d} " + " ".join(f"{rel[e][it]:10.0f}" if rel[e][it] is not None else " " * 10 for e in ORDER))
    inner = [i for i in range(4, n_it - 2) if all(abs(i - b) > 2 and abs(i - b - 1) > 2 for b in boundaries)]
    gaps = {}
    if len(inner) > 4:
        def mean_gap(a, b, shift=0):
            xs = [rel[b][i + shift] - rel[a][i] for i in inner if i + shift < ITERS and rel[a][i] is not None
                  and rel[b][i + shift] is not None]
            return sum(xs) / len(xs) if xs else float("nan")

        gaps = {"period (dS done -> dS done)": mean_gap(7, 7, 1),
                "S ready -> P done (P phase)": mean_gap(4, 5),
                "P done -> dP ready (compute idle)": mean_gap(5, 6),
                "dP ready -> dS done (dS phase)": mean_gap(6, 7),
                "dS done -> Y dK/dQ issue (handoff)": mean_gap(7, 3),
                "Y dK/dQ issue -> D dQ ready (dK+dQ MMA)": mean_gap(3, 8),
                "D dQ ready -> drained": mean_gap(8, 9),
                "D drained -> Y dP issue(next) (handoff)": mean_gap(9, 2, 1),
                "Y dP issue -> dP ready (dP MMA)": mean_gap(2, 6),
                "dS done -> S ready(next) (compute idle)": mean_gap(7, 4, 1)}
        for name, val in gaps.items():
            print(f"  {name:50s} {val:8.0f} ns")
    for b in boundaries[:6]:
        if b + 1 < n_it and rel[7][b + 1] is not None:
            print(f"  boundary after tile {b}: dS done -> dKV done {rel[15][b] - rel[7][b]:6.0f}, epilogue "
                  f"{rel[16][b] - rel[15][b]:6.0f}, dS done(last) -> dS done(first of next item) "
                  f"{rel[7][b + 1] - rel[7][b]:6.0f} ns; next item's K/V ready "
                  f"{(rel[17][b + 1] - rel[15][b]) if rel[17][b + 1] is not None else float('nan'):6.0f} ns after dKV done")
    if starts:
        first, last = min(starts), max(ends)
        busy = [e - s for s, e in zip(starts, ends)]
        print(f"CTA spans: {len(starts)} CTAs, launch span {(last - first) / 1e3:.1f} us, CTA busy mean "
              f"{sum(busy) / len(busy) / 1e3:.1f} us min {min(busy) / 1e3:.1f} max {max(busy) / 1e3:.1f}; "
              f"earliest finish {(min(ends) - first) / 1e3:.1f} us, latest {(last - first) / 1e3:.1f} us")
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"bwd_trace_n{n}_c{int(causal)}.json").write_text(json.dumps({"n": n, "causal": causal, "bh": bh,
                                                                          "gaps_ns": gaps, "rel_ns": rel}))


if __name__ == "__main__":
    main()
