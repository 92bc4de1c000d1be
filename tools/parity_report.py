"""Per-tensor parity report: max-abs, max-rel (|ref| > 1e-2), #violations at the reference tolerances
(5e-2/5e-2 for 16-bit O/dQ/dK/dV, 1e-3 for LSE) of the sm_100a kernels against the CPU oracle on the same inputs."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext
from oracle.attention_oracle import dense_backward_fp32, error_report

CASES = [  # bh, n, d, dtype, causal
    (2, 24, 64, "float16", True), (2, 32, 32, "float16", False), (4, 33, 64, "float16", True),
    (2, 512, 64, "bfloat16", True), (2, 512, 64, "bfloat16", False),      # C1 shape (16-bit)
    (2, 2048, 128, "bfloat16", True), (2, 2048, 128, "float16", True), (2, 4096, 128, "bfloat16", True),  # C2 slices
    (1, 8192, 128, "bfloat16", True), (2, 4096, 64, "bfloat16", False),
    (2, 333, 40, "bfloat16", True), (2, 1000, 96, "float16", False), (1, 16384, 128, "bfloat16", True),  # native d, long N
]
print("| bh | N | d | dtype | causal | tensor | max_abs | max_rel | violations / numel |")
print("|---|---|---|---|---|---|---|---|---|")
allrep = []
for bh, n, d, dt, causal in CASES:
    dtype = getattr(torch, dt)
    dpad = d if d % 8 == 0 else (64 if d <= 64 else 128)  # multiples of 8 run natively
    torch.manual_seed(0)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(4))
    scale = d ** -0.5
    if d == dpad:
        o, lse = ext.fwd_raw(q, k, v, causal, scale)
        dq, dk, dv = ext.bwd_raw(q, k, v, o, do, lse, causal, scale)
    else:
        o, lse = ext.forward(q, k, v, causal, scale, 128, 128)
        dq, dk, dv = ext.backward(q, k, v, o, do, lse, causal, scale, 128, 128)
    ref = dense_backward_fp32(*((q, k, v, do) if n > 4096 else (q.cpu(), k.cpu(), v.cpu(), do.cpu())), causal, scale)
    for name, got, want, tol in (("O", o, ref[3], 5e-2), ("LSE", lse, ref[4], 1e-3), ("dQ", dq, ref[0], 5e-2),
                                 ("dK", dk, ref[1], 5e-2), ("dV", dv, ref[2], 5e-2)):
        rep = error_report(got, want, tol, tol)
        allrep.append({"bh": bh, "n": n, "d": d, "dtype": dt, "causal": causal, "tensor": name, **rep})
        print(f"| {bh} | {n} | {d} | {dt} | {causal} | {name} | {rep['max_abs']:.3e} | {rep['max_rel']:.3e} | "
              f"{rep['violations']} / {rep['numel']} |", flush=True)
print("TOTAL_VIOLATIONS", sum(r["violations"] for r in allrep))
