"""torchrun script: ring attention over real NCCL vs the single-GPU kernel on the gathered tensors.
   torchrun --nproc-per-node P tools/ring_check.py [N] """
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import torch.distributed as dist
import flashattention_lab_cuda as ext
from dist.ring import contiguous_split, ring_attention, zigzag_split

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bh, d = 4, 128
worst = 0.0
for causal in (True, False):
    torch.manual_seed(123)  # same global tensors on every rank
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    scale = d ** -0.5
    o_ref, lse_ref = ext.fwd_raw(q, k, v, causal, scale)
    dq_ref, dk_ref, dv_ref = ext.bwd_raw(q, k, v, o_ref, do, lse_ref, causal, scale)
    split = zigzag_split if causal else contiguous_split
    ql, kl, vl = (split(t, world)[rank].requires_grad_(True) for t in (q, k, v))
    o, lse = ring_attention(ql, kl, vl, causal=causal, softmax_scale=scale)
    o.backward(split(do, world)[rank])
    torch.cuda.synchronize()
    errs = {}
    for name, got, ref in (("o", o, o_ref), ("dq", ql.grad, dq_ref), ("dk", kl.grad, dk_ref), ("dv", vl.grad, dv_ref)):
        errs[name] = (got.float() - split(ref, world)[rank].float()).abs().max().item()
    errs["lse"] = (lse - split(lse_ref.unsqueeze(-1), world)[rank].squeeze(-1)).abs().max().item()
    t = torch.tensor([max(errs.values())], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = max(worst, t.item())
    if rank == 0:
        print(f"ring_check world={world} N={n} causal={causal}: max|ring - single_gpu| per tensor =",
              {k_: f"{v_:.3e}" for k_, v_ in errs.items()}, flush=True)
if rank == 0:
    print("RING_CHECK", "OK" if worst < 3e-2 else "FAIL", f"worst={worst:.3e}", flush=True)
dist.destroy_process_group()
