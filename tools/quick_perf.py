"""Fast kernel-level check for iteration under gpurun: parity of fwd+bwd against the fp32 dense oracle on two small
shapes, then fwd / bwd-kernel TFLOP/s at C2 (N=4096 causal) and the headline shape (N=8192 non-causal and causal).
usage: [FA_SM100_LIB=...] python tools/quick_perf.py [--no-parity]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext


def parity():
    from oracle.attention_oracle import dense_backward_fp32, error_report
    bad = 0
    for (bh, n, d, causal, dt) in ((3, 640, 128, True, torch.bfloat16), (2, 1000, 128, False, torch.float16),
                                   (2, 384, 64, True, torch.bfloat16), (1, 2048, 128, True, torch.bfloat16)):
        torch.manual_seed(n)
        q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dt) for _ in range(4))
        o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
        dq, dk, dv = ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5)
        ref = dense_backward_fp32(q.cpu(), k.cpu(), v.cpu(), do.cpu(), causal, d ** -0.5)
        for name, got, want, tol in (("o", o, ref[3], 5e-2), ("lse", lse, ref[4], 1e-3), ("dq", dq, ref[0], 5e-2),
                                     ("dk", dk, ref[1], 5e-2), ("dv", dv, ref[2], 5e-2)):
            rep = error_report(got, want, tol, tol)
            if rep["violations"]:
                bad += 1
                print(f"PARITY FAIL bh={bh} n={n} d={d} causal={causal} {name}: {rep}", flush=True)
    print("parity:", "OK" if not bad else f"{bad} FAILURES", flush=True)
    return bad


def time_ms(fn, iters=10):
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


def perf():
    d = 128
    shapes = ((256, 1024, True), (128, 2048, True), (64, 4096, True), (64, 8192, False), (64, 8192, True))
    if "--short" in sys.argv:
        shapes = shapes[2:]
    for (bh, n, causal) in shapes:
        torch.manual_seed(0)
        q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
        o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
        flops_f = 4.0 * bh * n * n * d * (0.5 if causal else 1.0)
        t_f = time_ms(lambda: ext.fwd_raw(q, k, v, causal, d ** -0.5))
        if "--fwd-only" in sys.argv:
            print(f"bh={bh} n={n} causal={int(causal)}: fwd {t_f:.3f} ms {flops_f / t_f / 1e9:7.1f} TFLOP/s", flush=True)
            continue
        t_b = time_ms(lambda: ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5))
        acc = torch.empty(bh, n, d, device="cuda", dtype=torch.float32)
        stats = ext.bwd_prepare_raw(o, do, lse, zero=acc)
        t_m = time_ms(lambda: ext.bwd_raw(q, k, v, None, do, None, causal, d ** -0.5, rowstats=stats, dq_accum=acc))
        print(f"bh={bh} n={n} causal={int(causal)}: fwd {t_f:.3f} ms {flops_f / t_f / 1e9:7.1f} TFLOP/s | "
              f"bwd kernel {t_m:.3f} ms {2.5 * flops_f / t_m / 1e9:7.1f} | "
              f"bwd(all launches) {t_b:.3f} ms {2.5 * flops_f / t_b / 1e9:7.1f} TFLOP/s | "
              f"fwd+bwd {3.5 * flops_f / (t_f + t_b) / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    rc = 0 if "--no-parity" in sys.argv else parity()
    perf()
    sys.exit(1 if rc else 0)
