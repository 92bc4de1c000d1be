"""Small fixed workload for ncu: a few fwd (+bwd) launches at one shape.  usage: profile_target.py [fwd|bwd|both] B H N d causal"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext

what = sys.argv[1] if len(sys.argv) > 1 else "both"
b, h, n, d = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (4, 16, 4096, 128)
causal = (sys.argv[6] == "1") if len(sys.argv) > 6 else True
torch.manual_seed(0)
q, k, v, do = (torch.randn(b * h, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
scale = d ** -0.5
for _ in range(3):
    o, lse = ext.fwd_raw(q, k, v, causal, scale)
    if what in ("bwd", "both"):
        ext.bwd_raw(q, k, v, o, do, lse, causal, scale)
torch.cuda.synchronize()
print("done")
