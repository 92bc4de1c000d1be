import sys
sys.path[:0] = ["flashattention-pytorch_b200", "."]
import torch
import flashattention_lab_cuda as ext
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
out = []
for (bh, n, d, causal) in ((128, 4096, 64, True), (128, 8192, 64, False), (128, 8192, 64, True), (64, 8192, 128, False), (64, 8192, 128, True)):
    q, k, v = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    ms = t(lambda: ext.fwd_raw(q, k, v, causal, d ** -0.5))
    out.append(f"d{d} n{n} c{int(causal)}: {4.0*bh*n*n*d*(0.5 if causal else 1)/ms/1e9:.0f}")
print(" | ".join(out), flush=True)
