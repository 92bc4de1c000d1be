"""2-rank microbenchmark: bandwidth of the ring's block exchange primitives (NCCL batch_isend_irecv vs symmetric-memory
peer copies).  torchrun --nproc-per-node 2 tools/p2p_bw.py"""
import os
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nxt, prv = (rank + 1) % world, (rank - 1) % world


def bench(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for mb, ntens in ((64, 2), (256, 2), (256, 4)):
    sends = [torch.randn(mb * 2 ** 20 // 2, device=dev, dtype=torch.bfloat16) for _ in range(ntens)]
    recvs = [torch.empty_like(t) for t in sends]

    def exchange():
        ops = []
        for s, r in zip(sends, recvs):
            ops.append(dist.P2POp(dist.isend, s, nxt)); ops.append(dist.P2POp(dist.irecv, r, prv))
        for rq in dist.batch_isend_irecv(ops):
            rq.wait()

    ms = bench(exchange)
    if rank == 0:
        print(f"NCCL batch_isend_irecv {ntens} x {mb} MiB: {ms:.3f} ms -> {ntens * mb / 1024 / (ms * 1e-3):.0f} GiB/s per direction", flush=True)

try:
    import torch.distributed._symmetric_memory as symm
    n = 256 * 2 ** 20 // 2
    buf = symm.empty(n, dtype=torch.bfloat16, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    buf.normal_()
    dst = torch.empty(n, device=dev, dtype=torch.bfloat16)
    peer = hdl.get_buffer(prv, (n,), torch.bfloat16)
    hdl.barrier()

    def pull():
        dst.copy_(peer)

    ms = bench(pull)
    ok = bool(torch.isfinite(dst.float()).all())
    if rank == 0:
        print(f"symmetric-memory peer pull 256 MiB: {ms:.3f} ms -> {0.25 / (ms * 1e-3):.0f} GiB/s (copy engine), data ok={ok}", flush=True)
    hdl.barrier()
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("symmetric memory unavailable:", repr(e)[:300], flush=True)
dist.destroy_process_group()
