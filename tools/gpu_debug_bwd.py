import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext
torch.manual_seed(0)
bh, n, d = 1, 128, 128
q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = ext.fwd_raw(q, k, v, False, d ** -0.5)
torch.cuda.synchronize(); print("fwd ok", flush=True)
rs = ext.bwd_prepare_raw(o, do, lse)
torch.cuda.synchronize(); print("prepare ok", flush=True)
acc = torch.zeros(bh, n, d, device="cuda", dtype=torch.float32)
try:
    ext.bwd_raw(q, k, v, o, do, lse, False, d ** -0.5, rowstats=rs, dq_accum=acc)
    print("bwd launched", flush=True)
    torch.cuda.synchronize(); print("bwd ok", flush=True)
except Exception as e:
    print("EXC:", e, flush=True)
