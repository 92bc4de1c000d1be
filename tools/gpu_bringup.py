"""GPU bring-up script (run under gpurun): UMMA descriptor probe, then forward parity on a ladder of shapes.
Each stage runs in its own subprocess so a faulting kernel does not take the rest down."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]


def stage_probe(mode, dtype_name):
    import torch
    import probes

    dt = getattr(torch, dtype_name)
    torch.manual_seed(mode)
    a = torch.randn(256 if mode >= 4 else 128, 128, device="cuda", dtype=dt)
    b = torch.randn(128, 128, device="cuda", dtype=dt)
    out = probes.probe_umma(mode, a, b)
    torch.cuda.synchronize()
    af, bf = a.float(), b.float()
    want = {0: lambda: af @ bf.T, 1: lambda: af @ bf, 2: lambda: af @ bf, 3: lambda: af.T @ bf,
            4: lambda: af @ bf.T, 5: lambda: af @ bf}[mode]()
    err = (out - want).abs().max().item()
    print(f"probe mode={mode} {dtype_name}: max_abs_err={err:.4e} ref_max={want.abs().max().item():.2f}", flush=True)
    if err > 1e-2:
        # help diagnose layout mistakes: compare against a few plausible alternatives
        alts = {"A@B.T": af @ bf.T, "A@B": af @ bf}
        if mode < 4:
            alts.update({"A.T@B": af.T @ bf, "A.T@B.T": af.T @ bf.T})
        for name, alt in alts.items():
            print(f"    vs {name}: {(out - alt).abs().max().item():.4e}")
        print("    out[0,:8]", out[0, :8].tolist())
        print("    want[0,:8]", want[0, :8].tolist())
        return 1
    return 0


def stage_fwd(bh, n, d, dtype_name, causal):
    import torch
    import flashattention_lab_cuda as ext
    from oracle.attention_oracle import dense_forward, error_report

    dt = getattr(torch, dtype_name)
    torch.manual_seed(1234)
    q, k, v = (torch.randn(bh, n, d, device="cuda", dtype=dt) for _ in range(3))
    o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    torch.cuda.synchronize()
    o_ref, lse_ref = dense_forward(q.cpu(), k.cpu(), v.cpu(), causal, d ** -0.5)
    ro = error_report(o, o_ref, 5e-2, 5e-2)
    rl = error_report(lse, lse_ref, 1e-3, 1e-3)
    ok = ro["violations"] == 0 and rl["violations"] == 0
    print(f"fwd bh={bh} n={n} d={d} {dtype_name} causal={causal}: O max_abs={ro['max_abs']:.3e} viol={ro['violations']} "
          f"| LSE max_abs={rl['max_abs']:.3e} viol={rl['violations']} -> {'OK' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


def stage_bwd(bh, n, d, dtype_name, causal):
    import torch
    import flashattention_lab_cuda as ext
    from oracle.attention_oracle import dense_backward_fp32, error_report

    dt = getattr(torch, dtype_name)
    torch.manual_seed(4321)
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=dt) for _ in range(4))
    o, lse = ext.fwd_raw(q, k, v, causal, d ** -0.5)
    dq, dk, dv = ext.bwd_raw(q, k, v, o, do, lse, causal, d ** -0.5)
    torch.cuda.synchronize()
    dq_r, dk_r, dv_r, _, _ = dense_backward_fp32(q.cpu(), k.cpu(), v.cpu(), do.cpu(), causal, d ** -0.5)
    ok = True
    msg = []
    for name, got, want in (("dQ", dq, dq_r), ("dK", dk, dk_r), ("dV", dv, dv_r)):
        rep = error_report(got, want, 5e-2, 5e-2)
        ok &= rep["violations"] == 0
        msg.append(f"{name} max_abs={rep['max_abs']:.3e} viol={rep['violations']}")
    print(f"bwd bh={bh} n={n} d={d} {dtype_name} causal={causal}: " + " | ".join(msg) + f" -> {'OK' if ok else 'FAIL'}",
          flush=True)
    return 0 if ok else 1


def stage_perf(b, h, n, d, causal):
    import torch
    import flashattention_lab_cuda as ext

    torch.manual_seed(0)
    bh = b * h
    q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    scale = d ** -0.5
    o, lse = ext.fwd_raw(q, k, v, causal, scale)
    c = 0.5 if causal else 1.0
    f_fwd = 4 * bh * n * n * d * c
    f_bwd = 10 * bh * n * n * d * c

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    t_f = timeit(lambda: ext.fwd_raw(q, k, v, causal, scale, out=o, lse=lse))
    line = f"perf B={b} H={h} N={n} d={d} causal={causal}: fwd {t_f:.3f} ms = {f_fwd / t_f / 1e9:.0f} TFLOP/s"
    try:
        t_b = timeit(lambda: ext.bwd_raw(q, k, v, o, do, lse, causal, scale))
        line += f" | bwd(total) {t_b:.3f} ms = {f_bwd / t_b / 1e9:.0f} TFLOP/s"
    except Exception as exc:  # noqa: BLE001
        line += f" | bwd failed: {exc}"
    print(line, flush=True)
    return 0


def main():
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        kind = sys.argv[1]
        if kind == "probe":
            sys.exit(stage_probe(int(sys.argv[2]), sys.argv[3]))
        if kind == "fwd":
            sys.exit(stage_fwd(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6] == "1"))
        if kind == "bwd":
            sys.exit(stage_bwd(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6] == "1"))
        if kind == "perf":
            sys.exit(stage_perf(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6] == "1"))
    stages = [["probe", str(m), dt] for m in range(4) for dt in ("bfloat16", "float16")]
    for bh, n, d, dt in [(1, 128, 128, "bfloat16"), (2, 256, 128, "bfloat16"), (2, 33, 64, "float16"),
                         (3, 300, 64, "float16"), (2, 1024, 128, "bfloat16"), (2, 777, 128, "float16"),
                         (4, 2048, 64, "bfloat16"), (8, 4096, 128, "bfloat16"), (150, 1536, 128, "float16"),
                         (2, 8192, 64, "float16")]:
        for causal in ("0", "1"):
            stages.append(["fwd", str(bh), str(n), str(d), dt, causal])
    if "--no-probe" in sys.argv or True:
        stages = [st for st in stages if st[0] != "probe"] if "--skip-probe" in sys.argv else stages
    for bh, n, d, dt in [(1, 128, 128, "bfloat16"), (2, 256, 128, "bfloat16"), (2, 33, 64, "float16"),
                         (3, 300, 64, "float16"), (2, 1024, 128, "bfloat16"), (2, 777, 128, "float16"),
                         (4, 2048, 64, "bfloat16")]:
        for causal in ("0", "1"):
            stages.append(["bwd", str(bh), str(n), str(d), dt, causal])
    for b, h, n, d in [(4, 16, 4096, 128), (4, 16, 8192, 128), (4, 32, 4096, 64)]:
        for causal in ("1", "0"):
            stages.append(["perf", str(b), str(h), str(n), str(d), causal])
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
    if only:
        stages = [st for st in stages if st[0] in only[0]]
    fails = 0
    for st in stages:
        try:
            r = subprocess.run([sys.executable, __file__, *st], timeout=180)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
        if rc != 0:
            fails += 1
            print(f"STAGE {' '.join(st)} -> rc={rc}", flush=True)
    print(f"bring-up done: {fails} failing stages of {len(stages)}")
    sys.exit(0)


if __name__ == "__main__":
    main()
