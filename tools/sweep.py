"""Benchmark sweep in the reference's harness conventions: the CLI flags of
`benchmarks/bench_utils.py:250-264` (add_common_args), records in the `BenchmarkRecord` schema (`:161-180`) written as
JSON + CSV with the field order of `write_results` (`:287-325`), so the reference's plotting loads them unchanged —
over the FA1 / FA2 / FA3 entry points like `benchmarks/bench_compare_all.py`.

    python tools/sweep.py --c3                     # BASELINE config C3: fa1+fa3, N 1K..16K, d 64/128, fp16/bf16,
                                                   # constant 16K tokens and hidden 2048 (B = 16384/N, H = 2048/d)
    python tools/sweep.py --algos fa1 fa2 fa3 --seqlen 512 1024 --head-dim 64 128 256 --batch-size 1 2 --num-heads 4

Timing: CUDA events around every call (the reference uses perf_counter + synchronize, `:128-132`), `--warmup` untimed
and `--iters` timed calls per point; inputs as `bench_utils.py:83-97` (seed 0; q, k, v, then dO from one generator).
`tflops` follows the reference's convention (forward 4*B*H*N^2*D, "backward" = fwd+bwd time with 8*B*H*N^2*D, no
causal discount, `:210-215`); the algorithmic rate (14*B*H*N^2*D*(1/2 if causal) for fwd+bwd) goes into `config`."""
import argparse
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]

# reference benchmarks/bench_utils.py:161-180 (dataclass fields) == the CSV header of write_results (:300-319)
REF_FIELDS = ["method", "algo", "backend", "direction", "dtype", "causal", "seqlen", "head_dim", "batch_size",
              "num_heads", "mean_ms", "std_ms", "tflops", "peak_mem_mb", "status", "fp8", "config", "error"]
NOMINAL = 2250.0
DTYPES = {"fp16": "float16", "float16": "float16", "bf16": "bfloat16", "bfloat16": "bfloat16", "fp32": "float32",
          "float32": "float32"}


def add_common_args(parser):  # reference benchmarks/bench_utils.py:250-264, same flags and defaults
    parser.add_argument("--device", default="cuda", choices=["cpu", "cuda"])
    parser.add_argument("--seqlen", type=int, nargs="+", default=[512, 1024, 2048, 4096, 8192, 16384])
    parser.add_argument("--head-dim", type=int, nargs="+", default=[64, 128, 256])
    parser.add_argument("--batch-size", type=int, nargs="+", default=[1, 2])
    parser.add_argument("--num-heads", type=int, nargs="+", default=[4])
    parser.add_argument("--causal", action="store_true", help="Run only causal mode (defaults to both)")
    parser.add_argument("--non-causal-only", action="store_true", help="Run only non-causal mode")
    parser.add_argument("--dtypes", type=str, nargs="+", default=["fp16", "bf16"], help="Dtypes: fp16 bf16 fp32")
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--iters", type=int, default=20)


def iter_causal_flags(args):  # reference benchmarks/bench_utils.py:267-272
    if args.causal:
        return [True]
    if args.non_causal_only:
        return [False]
    return [False, True]


def make_record(algo, direction, dtype, causal, n, d, b, h, mean_ms, std_ms, peak_mb, status="ok", error=None,
                algorithmic_tflops=None, fp8=False):
    """One row in the reference's BenchmarkRecord schema (method label with an " FP8" suffix as
    benchmarks/bench_compare_all.py:65-67)."""
    factor = 4.0 if direction == "forward" else 8.0  # reference attention_flops (:210-215)
    tflops = None if mean_ms is None else factor * b * h * n * n * d / (mean_ms * 1e-3) / 1e12
    return {"method": f"{algo.upper()} (sm_100a)" + (" FP8" if fp8 else ""), "algo": algo, "backend": "cuda", "direction": direction,
            "dtype": dtype, "causal": bool(causal), "seqlen": n, "head_dim": d, "batch_size": b, "num_heads": h,
            "mean_ms": mean_ms, "std_ms": std_ms, "tflops": tflops, "peak_mem_mb": peak_mb, "status": status,
            "fp8": bool(fp8) if algo == "fa3" else None,
            "config": None if algorithmic_tflops is None else f"algorithmic_tflops={algorithmic_tflops:.1f}",
            "error": error}


def write_results(name, records, out_dir):
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / f"{name}.json").write_text(json.dumps(records, indent=2))
    with (out_dir / f"{name}.csv").open("w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=REF_FIELDS)
        w.writeheader()
        w.writerows(records)


def timeit(fn, warmup, iters):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    mean = sum(ts) / len(ts)
    return mean, (sum((t - mean) ** 2 for t in ts) / len(ts)) ** 0.5


def run_point(api_name, api, dtype_name, d, n, b, h, causal, warmup, iters, fp8=False):
    import torch

    extra = {"fp8": True} if fp8 else {}  # FA3 only: e4m3 forward, 16-bit straight-through backward

    dtype = getattr(torch, DTYPES[dtype_name])
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn((b, h, n, d), generator=g, device="cuda", dtype=dtype).requires_grad_(True) for _ in range(3))
    do = torch.randn((b, h, n, d), generator=g, device="cuda", dtype=dtype)
    f_fwd = 4.0 * b * h * n * n * d * (0.5 if causal else 1.0)
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad():
        t_f, sd_f = timeit(lambda: api(q, k, v, causal=causal, backend="cuda", **extra), warmup, iters)
    o, _ = api(q, k, v, causal=causal, backend="cuda", **extra)

    def bwd():
        torch.autograd.backward(o, do, retain_graph=True)
        q.grad = k.grad = v.grad = None

    try:
        t_b, sd_b = timeit(bwd, warmup, iters)
    except NotImplementedError as exc:  # e.g. head dim 256: the forward exists, the backward does not
        peak_mb = torch.cuda.max_memory_allocated() / 2 ** 20
        row = {"api": api_name, "dtype": dtype_name, "d": d, "N": n, "B": b, "H": h, "causal": causal, "fwd_ms": t_f,
               "bwd_ms": None, "fwd_tflops": f_fwd / t_f / 1e9, "bwd_tflops": None, "fwd_bwd_tflops": None,
               "frac_nominal": None}
        recs = [make_record(api_name, "forward", dtype_name, causal, n, d, b, h, t_f, sd_f, peak_mb,
                            algorithmic_tflops=row["fwd_tflops"], fp8=fp8),
                make_record(api_name, "backward", dtype_name, causal, n, d, b, h, None, None, None, "unsupported",
                            str(exc)[:200], fp8=fp8)]
        return row, recs
    peak_mb = torch.cuda.max_memory_allocated() / 2 ** 20
    row = {"api": api_name, "dtype": dtype_name, "d": d, "N": n, "B": b, "H": h, "causal": causal, "fwd_ms": t_f,
           "bwd_ms": t_b, "fwd_tflops": f_fwd / t_f / 1e9, "bwd_tflops": 2.5 * f_fwd / t_b / 1e9,
           "fwd_bwd_tflops": 3.5 * f_fwd / (t_f + t_b) / 1e9}
    row["frac_nominal"] = row["fwd_bwd_tflops"] / NOMINAL
    recs = [make_record(api_name, "forward", dtype_name, causal, n, d, b, h, t_f, sd_f, peak_mb,
                        algorithmic_tflops=row["fwd_tflops"], fp8=fp8),
            make_record(api_name, "backward", dtype_name, causal, n, d, b, h, t_f + t_b, (sd_f ** 2 + sd_b ** 2) ** 0.5,
                        peak_mb, algorithmic_tflops=row["fwd_bwd_tflops"], fp8=fp8)]
    return row, recs


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    add_common_args(ap)
    ap.add_argument("--algos", nargs="+", default=["fa1", "fa2", "fa3"], choices=["fa1", "fa2", "fa3"])
    ap.add_argument("--c3", action="store_true", help="BASELINE config C3 preset (overrides shapes / dtypes / algos)")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out"), help="directory for <name>.json / <name>.csv")
    ap.add_argument("--name", "--tag", dest="name", default="sweep_records",
                    help="base name of the result files (--tag: reference benchmarks/bench_compare_all.py:81)")
    ap.add_argument("--fp8", action="store_true",
                    help="also run FA3 with fp8=True at every point (reference benchmarks/bench_compare_all.py:73)")
    ap.add_argument("--directions", nargs="+", default=["forward", "backward"], choices=["forward", "backward"],
                    help="which records to write (both are always timed; reference bench_compare_all.py:74-80)")
    args = ap.parse_args()
    import torch
    from fa1 import fa1_attention
    from fa2 import fa2_attention
    from fa3 import fa3_attention

    apis = {"fa1": fa1_attention, "fa2": fa2_attention, "fa3": fa3_attention}
    if args.c3:
        points = [(a, dt, d, n, 16384 // n, 2048 // d, c, False) for a in ("fa1", "fa3") for dt in ("bf16", "fp16")
                  for d in (128, 64) for n in (1024, 2048, 4096, 8192, 16384) for c in (True, False)]
        warmup, iters = 3, 10
    else:
        points = [(a, dt, d, n, b, h, c, f8) for a in args.algos for dt in args.dtypes for d in args.head_dim
                  for n in args.seqlen for b in args.batch_size for h in args.num_heads for c in iter_causal_flags(args)
                  for f8 in ([False, True] if (a == "fa3" and args.fp8) else [False])]
        warmup, iters = args.warmup, args.iters
    rows, records = [], []
    print("| api | dtype | d | N | B | H | causal | fwd ms | fwd TF/s | bwd ms | bwd TF/s | fwd+bwd TF/s | % of 2250 |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for api_name, dt, d, n, b, h, causal, fp8 in points:
        label = api_name + ("+fp8" if fp8 else "")
        try:
            row, recs = run_point(api_name, apis[api_name], dt, d, n, b, h, causal, warmup, iters, fp8=fp8)
        except (NotImplementedError, RuntimeError) as exc:  # unsupported point: recorded like the reference does
            oom = "out of memory" in str(exc).lower()
            status = "oom" if oom else "unsupported"
            records += [make_record(api_name, direction, dt, causal, n, d, b, h, None, None, None, status, str(exc)[:200],
                                    fp8=fp8) for direction in args.directions]
            print(f"| {label} | {dt} | {d} | {n} | {b} | {h} | {causal} | - | - | - | - | {status} | - |", flush=True)
            torch.cuda.empty_cache()
            continue
        row["api"] = label
        rows.append(row)
        records += [r for r in recs if r["direction"] in args.directions]
        if row["bwd_ms"] is None:
            print(f"| {label} | {dt} | {d} | {n} | {b} | {h} | {causal} | {row['fwd_ms']:.3f} | {row['fwd_tflops']:.0f} | "
                  f"- | - | forward only | - |", flush=True)
            continue
        print(f"| {label} | {dt} | {d} | {n} | {b} | {h} | {causal} | {row['fwd_ms']:.3f} | {row['fwd_tflops']:.0f} | "
              f"{row['bwd_ms']:.3f} | {row['bwd_tflops']:.0f} | {row['fwd_bwd_tflops']:.0f} | "
              f"{100 * row['frac_nominal']:.1f} |", flush=True)
    print("JSON " + json.dumps(rows))
    write_results(args.name, records, Path(args.out))


if __name__ == "__main__":
    main()
