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


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


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
                    with torch.no_grad():
                        t_f = timeit(lambda: api(q, k, v, causal=causal, backend="cuda"))
                    o, _ = api(q, k, v, causal=causal, backend="cuda")

                    def bwd():
                        torch.autograd.backward(o, do, retain_graph=True)
                        q.grad = k.grad = v.grad = None

                    t_b = timeit(bwd)
                    tf_f, tf_b = f_fwd / t_f / 1e9, 2.5 * f_fwd / t_b / 1e9
                    tf_fb = 3.5 * f_fwd / (t_f + t_b) / 1e9
                    rec = {"api": api_name, "dtype": str(dtype).split(".")[-1], "d": d, "N": n, "B": b, "H": h,
                           "causal": causal, "fwd_ms": t_f, "bwd_ms": t_b, "fwd_tflops": tf_f, "bwd_tflops": tf_b,
                           "fwd_bwd_tflops": tf_fb, "frac_nominal": tf_fb / NOMINAL}
                    rows.append(rec)
                    print(f"| {api_name} | {rec['dtype']} | {d} | {n} | {causal} | {t_f:.3f} | {tf_f:.0f} | {t_b:.3f} | "
                          f"{tf_b:.0f} | {tf_fb:.0f} | {100 * tf_fb / NOMINAL:.1f} |", flush=True)
print("JSON " + json.dumps(rows))
