"""CTA-pair (cta_group::2) bring-up and tensor-core issue-rate sweep (run under gpurun).

Stage 1: fa_sm100_probe_umma modes 4/5 against torch matmul (each in its own subprocess, with a timeout, so a wrong
descriptor traps or times out without taking the rest down).
Stage 2: fa_sm100_probe_mma_rate over the operand configurations the attention kernels use or could use; prints
TFLOP/s for the whole GPU and per-product time per SM.  Results are appended to gpurun_out/pair_probe.json.
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]


def stage_rate(mma=True):
    import torch
    import probes

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ctas = sms - (sms & 1)
    groups = 4096
    rows = []
    configs = [  # (pair, a_from_tmem, n)
        (0, 0, 64), (0, 0, 128), (0, 0, 256), (0, 1, 64), (0, 1, 128),
        (1, 0, 64), (1, 0, 128), (1, 0, 256), (1, 1, 128),
    ]
    for pair, ts, n in (configs if mma else []):
        for _ in range(2):
            probes.probe_mma_rate(pair, ts, n, 256, ctas)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            probes.probe_mma_rate(pair, ts, n, groups, ctas)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        flops = ctas * groups * 2.0 * 128 * n * 128
        row = {"pair": pair, "a_from_tmem": ts, "n": n, "ctas": ctas, "groups": groups, "ms": best,
               "tflops": flops / best / 1e9, "ns_per_product_per_sm": best * 1e6 / groups}
        rows.append(row)
        print(f"rate pair={pair} ts={ts} n={n:3d}: {row['tflops']:8.1f} TFLOP/s  "
              f"{row['ns_per_product_per_sm']:7.1f} ns per (128 x {n} x 128) product per SM", flush=True)
    # L2 reduce-add rate with the backward's dQ pattern (headline shape: 64 slices, N = 8192 -> 64 x 64 tiles)
    for slices, nqt, nkt in ((64, 64, 64), (64, 32, 32)):
        acc = torch.zeros(slices, nqt * 128, 128, device="cuda", dtype=torch.float32)
        n_calls = 0
        for flags in (0, 1, 4, 2, 3, 6):
            probes.probe_reduce_rate(acc, nkt, flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            probes.probe_reduce_rate(acc, nkt, flags)
            e1.record()
            torch.cuda.synchronize()
            n_calls += 2
            ms = e0.elapsed_time(e1)
            gb = slices * nkt * nqt * 65536 / 1e9
            row = {"probe": "reduce_rate", "slices": slices, "nqt": nqt, "nkt": nkt, "flags": flags, "ms": ms,
                   "gbytes": gb, "tb_per_s": gb / ms}
            rows.append(row)
            print(f"reduce slices={slices} nqt={nqt} nkt={nkt} flags={flags} "
                  f"({'regs' if flags & 2 else 'tma '}{' rot' if flags & 1 else ''}{' 1cta/sm' if flags & 4 else ''}): "
                  f"{gb:.1f} GB in {ms:.3f} ms = {gb / ms:.2f} TB/s", flush=True)
        want = float(n_calls * nkt)
        ok = bool((acc == want).all().item())
        print(f"  accumulated value check (every element == {want}): {ok}", flush=True)
        del acc
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "pair_probe.json").write_text(json.dumps(rows, indent=1))
    return 0


def main():
    if len(sys.argv) > 1 and sys.argv[1] in ("rate", "reduce"):
        sys.exit(stage_rate(mma=sys.argv[1] == "rate"))
    rc = 0
    for mode in (4, 5):
        for dt in ("bfloat16", "float16"):
            try:
                r = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_bringup.py"), "probe", str(mode), dt],
                                   timeout=120)
                rc |= r.returncode != 0
            except subprocess.TimeoutExpired:
                print(f"probe mode={mode} {dt}: TIMEOUT", flush=True)
                rc |= 1
    try:
        r = subprocess.run([sys.executable, __file__, "rate"], timeout=300)
        rc |= r.returncode != 0
    except subprocess.TimeoutExpired:
        print("rate sweep: TIMEOUT", flush=True)
        rc |= 1
    sys.exit(int(rc))


if __name__ == "__main__":
    main()
