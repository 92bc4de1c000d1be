"""Workload for compute-sanitizer (tools/gpu_sanitize.sh): forward + backward (+ the ring forms) at four small shapes —
causal / ragged / both head-dim variants / both dtypes — through the C ABI, each checked against the fp32 oracle."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext
from oracle.attention_oracle import dense_backward_fp32, error_report

SHAPES = [(2, 384, 128, torch.bfloat16, True), (1, 300, 64, torch.float16, True), (2, 257, 96, torch.bfloat16, False),
          (1, 130, 40, torch.float16, True)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for idx, (bh, n, d, dtype, causal) in enumerate(SHAPES):
    if only is not None and idx != only:
        continue
    torch.manual_seed(idx)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(4))
    scale = d ** -0.5
    o, lse = ext.fwd_raw(q, k, v, causal, scale)
    dq, dk, dv = ext.bwd_raw(q, k, v, o, do, lse, causal, scale)
    acc = torch.empty(bh, n, d, device="cuda", dtype=torch.float32)
    stats = ext.bwd_prepare_raw(o, do, lse, zero=acc)
    dk_acc, dv_acc = (torch.zeros(bh, n, d, device="cuda", dtype=torch.float32) for _ in range(2))
    ext.bwd_raw(q, k, v, None, do, None, causal, scale, rowstats=stats, dq_accum=acc, dk_accum=dk_acc, dv_accum=dv_acc)
    ext.fwd_raw(q, k, v, causal, scale, out=o.clone(), lse=lse.clone(), merge=True)
    torch.cuda.synchronize()
    dq_r, dk_r, dv_r, o_r, lse_r = dense_backward_fp32(q.cpu(), k.cpu(), v.cpu(), do.cpu(), causal, scale)
    bad = sum(error_report(a, b, t, t)["violations"] for a, b, t in
              ((o, o_r, 5e-2), (lse, lse_r, 1e-3), (dq, dq_r, 5e-2), (dk, dk_r, 5e-2), (dv, dv_r, 5e-2),
               (dk_acc, dk_r, 5e-2), (dv_acc, dv_r, 5e-2)))
    print(f"shape {idx} bh={bh} n={n} d={d} {dtype} causal={causal}: violations={bad}", flush=True)
    assert bad == 0
print("SANITIZE_TARGET_OK")
