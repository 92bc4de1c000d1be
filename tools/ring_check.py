"""torchrun script: ring attention over real ranks (NCCL / symmetric memory) vs (a) the single-GPU kernel on the
gathered tensors and (b) the dense fp32 oracle (test infrastructure, N <= 16384), per tensor.
   torchrun --nproc-per-node P tools/ring_check.py [N] [--no-oracle]
Prints `RING_CHECK OK|FAIL ...`; exit code 1 on FAIL (tests/test_ring_multigpu.py runs it)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import torch.distributed as dist
import flashattention_lab_cuda as ext
from dist.ring import contiguous_split, ring_attention, ring_transport_name, zigzag_split

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 8192
use_oracle = "--no-oracle" not in sys.argv and n <= 16384
bh, d = 4, 128
worst, worst_oracle = 0.0, 0.0
for causal in (True, False):
    torch.manual_seed(123)  # same global tensors on every rank
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    scale = d ** -0.5
    o_ref, lse_ref = ext.fwd_raw(q, k, v, causal, scale)
    dq_ref, dk_ref, dv_ref = ext.bwd_raw(q, k, v, o_ref, do, lse_ref, causal, scale)
    split = zigzag_split if causal else contiguous_split
    mine = lambda t: split(t, world)[rank]  # noqa: E731
    ql, kl, vl = (mine(t).requires_grad_(True) for t in (q, k, v))
    o, lse = ring_attention(ql, kl, vl, causal=causal, softmax_scale=scale)
    o.backward(mine(do))
    torch.cuda.synchronize()
    errs = {}
    for name, got, ref in (("o", o, o_ref), ("dq", ql.grad, dq_ref), ("dk", kl.grad, dk_ref), ("dv", vl.grad, dv_ref)):
        errs[name] = (got.float() - mine(ref).float()).abs().max().item()
    errs["lse"] = (lse - mine(lse_ref.unsqueeze(-1)).squeeze(-1)).abs().max().item()
    t = torch.tensor([max(errs.values())], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = max(worst, t.item())
    if rank == 0:
        print(f"ring_check world={world} N={n} causal={causal} transport={ring_transport_name(ql)}: "
              f"max|ring - single_gpu| per tensor =", {k_: f"{v_:.3e}" for k_, v_ in errs.items()}, flush=True)
    if use_oracle:
        from oracle.attention_oracle import dense_backward_fp32  # the fp32 restatement, run on this GPU slice by slice

        oerr = {}
        for s_ in range(bh):
            sl = slice(s_, s_ + 1)
            dq_o, dk_o, dv_o, o_o, lse_o = dense_backward_fp32(q[sl], k[sl], v[sl], do[sl], causal, scale)
            for name, got, ref in (("o", o[sl], o_o), ("dq", ql.grad[sl], dq_o), ("dk", kl.grad[sl], dk_o),
                                   ("dv", vl.grad[sl], dv_o)):
                oerr[name] = max(oerr.get(name, 0.0), (got.float() - mine(ref)).abs().max().item())
            oerr["lse"] = max(oerr.get("lse", 0.0), (lse[sl] - mine(lse_o.unsqueeze(-1)).squeeze(-1)).abs().max().item())
        lse_bad = oerr.pop("lse")
        t = torch.tensor([max(oerr.values()), lse_bad], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst_oracle = max(worst_oracle, t[0].item(), t[1].item() * 50)  # lse tolerance 1e-3 vs 5e-2
        if rank == 0:
            print(f"ring_check world={world} N={n} causal={causal}: max|ring - fp32 oracle| =",
                  {k_: f"{v_:.3e}" for k_, v_ in oerr.items()}, f"lse {lse_bad:.3e}", flush=True)
ok = worst < 3e-2 and worst_oracle < 5e-2
if rank == 0:
    print("RING_CHECK", "OK" if ok else "FAIL", f"worst_vs_kernel={worst:.3e} worst_vs_oracle={worst_oracle:.3e}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
