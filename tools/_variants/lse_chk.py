import sys
sys.path[:0] = ["flashattention-pytorch_b200", "."]
import torch
import flashattention_lab_cuda as ext
from oracle.attention_oracle import dense_forward
torch.manual_seed(0)
for d in (128, 64):
    q, k, v = (torch.randn(2, 384, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    o, lse = ext.fwd_raw(q, k, v, True, d ** -0.5)
    o_r, lse_r = dense_forward(q.cpu(), k.cpu(), v.cpu(), True, d ** -0.5)
    print(f"d={d}: lse max err {(lse.cpu() - lse_r).abs().max().item():.3e}  o max err {(o.float().cpu() - o_r.float()).abs().max().item():.3e}  lse checksum {lse.double().sum().item():.10f}")
