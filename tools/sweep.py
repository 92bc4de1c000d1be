"""BASELINE config C3: FA1 and FA3 entry points, fwd / bwd / fwd+bwd sweep over N in {1K..16K}, d in {64,128},
fp16/bf16, causal and non-causal, constant 16K tokens and hidden size 2048 (B = 16384/N, H = 2048/d), as the
reference's benchmark conventions (benchmarks/bench_utils.py:83-97: seed 0, q,k,v,dO = randn) prescribe.
Prints a markdown table + one JSON line per point.  Timing: CUDA events, 3 warm-up + 10 timed calls per point."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
from fa1 import fa1_attention
from fa3 import fa3_attention

NOMINAL = 2250.0


def timeit(fn, iters=10, stats=None):
    """mean ms per call; per-call CUDA events so a std can be reported like the reference's benchmark_fn."""
    for _ in range(3):
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
    if stats is not None:
        stats["std"] = (sum((t - mean) ** 2 for t in ts) / len(ts)) ** 0.5
    return mean


# Records in the reference's own schema (benchmarks/bench_utils.py:161-207 BenchmarkRecord; :210-215 FLOP convention:
# forward 4*B*H*N^2*D, "backward" 8*B*H*N^2*D for the fwd+bwd time, NO causal discount) so its plotting can load them.
REF_FIELDS = ["method", "algo", "backend", "direction", "dtype", "causal", "seqlen", "head_dim", "batch_size",
              "num_heads", "mean_ms", "std_ms", "tflops", "peak_mem_mb", "status", "fp8", "config", "error"]
ref_records = []


def ref_record(api_name, direction, dtype, causal, n, d, b, h, mean_ms, std_ms, peak_mb, algo_tflops):
    factor = 4.0 if direction == "forward" else 8.0
    ref_records.append({
        "method": f"{api_name.upper()} (sm_100a)", "algo": api_name, "backend": "cuda", "direction": direction,
        "dtype": dtype, "causal": bool(causal), "seqlen": n, "head_dim": d, "batch_size": b, "num_heads": h,
        "mean_ms": mean_ms, "std_ms": std_ms, "tflops": factor * b * h * n * n * d / (mean_ms * 1e-3) / 1e12,
        "peak_mem_mb": peak_mb, "status": "ok", "fp8": False if api_name == "fa3" else None,
        "config": f"algorithmic_tflops={algo_tflops:.1f}", "error": None})


rows = []
print("| api | dtype | d | N | causal | fwd ms | fwd TF/s | bwd ms | bwd TF/s | fwd+bwd TF/s | % of 2250 |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for api_name, api in (("fa1", fa1_attention), ("fa3", fa3_attention)):
    for dtype in (torch.bfloat16, torch.float16):
        for d in (128, 64):
            for n in (1024, 2048, 4096, 8192, 16384):
                for causal in (True, False):
                    b, h = 16384 // n, 2048 // d
                    g = torch.Generator(device="cuda").manual_seed(0)
                    q, k, v = (torch.randn((b, h, n, d), generator=g, device="cuda", dtype=dtype).requires_grad_(True)
                               for _ in range(3))
                    do = torch.randn((b, h, n, d), generator=g, device="cuda", dtype=dtype)
                    c = 0.5 if causal else 1.0
                    f_fwd = 4.0 * b * h * n * n * d * c
                    torch.cuda.reset_peak_memory_stats()
                    st_f, st_b = {}, {}
                    with torch.no_grad():
                        t_f = timeit(lambda: api(q, k, v, causal=causal, backend="cuda"), stats=st_f)
                    o, _ = api(q, k, v, causal=causal, backend="cuda")

                    def bwd():
                        torch.autograd.backward(o, do, retain_graph=True)
                        q.grad = k.grad = v.grad = None

                    t_b = timeit(bwd, stats=st_b)
                    peak_mb = torch.cuda.max_memory_allocated() / 2 ** 20
                    tf_f, tf_b = f_fwd / t_f / 1e9, 2.5 * f_fwd / t_b / 1e9
                    tf_fb = 3.5 * f_fwd / (t_f + t_b) / 1e9
                    rec = {"api": api_name, "dtype": str(dtype).split(".")[-1], "d": d, "N": n, "B": b, "H": h,
                           "causal": causal, "fwd_ms": t_f, "bwd_ms": t_b, "fwd_tflops": tf_f, "bwd_tflops": tf_b,
                           "fwd_bwd_tflops": tf_fb, "frac_nominal": tf_fb / NOMINAL}
                    rows.append(rec)
                    dn = str(dtype).split(".")[-1]
                    ref_record(api_name, "forward", dn, causal, n, d, b, h, t_f, st_f["std"], peak_mb, f_fwd / t_f / 1e9)
                    ref_record(api_name, "backward", dn, causal, n, d, b, h, t_f + t_b, (st_f["std"] ** 2 + st_b["std"] ** 2) ** 0.5,
                               peak_mb, 3.5 * f_fwd / (t_f + t_b) / 1e9)
                    print(f"| {api_name} | {rec['dtype']} | {d} | {n} | {causal} | {t_f:.3f} | {tf_f:.0f} | {t_b:.3f} | "
                          f"{tf_b:.0f} | {tf_fb:.0f} | {100 * tf_fb / NOMINAL:.1f} |", flush=True)
print("JSON " + json.dumps(rows))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / "sweep_records.json").write_text(json.dumps(ref_records, indent=2))
import csv
with (out / "sweep_records.csv").open("w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=REF_FIELDS)
    w.writeheader()
    w.writerows(ref_records)
