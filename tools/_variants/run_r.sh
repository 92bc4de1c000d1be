timeout 600 python -m pytest tests/test_kernel_parity.py -m gpu -x -q -k "head_dims_up_to_256 or shim" 2>&1 | tail -25
python - <<'PY'
import sys, torch
sys.path[:0] = ["flashattention-pytorch_b200", "."]
import flashattention_lab_cuda as ext
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
for (bh, n, d, causal) in ((32, 4096, 256, True), (32, 8192, 256, False), (32, 8192, 256, True)):
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    f = 4.0 * bh * n * n * d * (0.5 if causal else 1)
    tf = t(lambda: ext.fwd_raw(q, k, v, causal, d ** -0.5)); tb = t(lambda: ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5))
    print(f"d={d} bh={bh} n={n} causal={causal}: fwd {tf:.3f} ms {f/tf/1e9:.0f} TFLOP/s, bwd {tb:.3f} ms {2.5*f/tb/1e9:.0f} TFLOP/s", flush=True)
PY
