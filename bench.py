#!/usr/bin/env python
"""bench.py — attention fwd+bwd throughput on B200 (the metric of BASELINE.json) and the reference CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|headline|c4|c5]

One "step" = one forward + one backward of `fa2_attention` (the reference's public entry point) over one batch of
synthetic Q/K/V/dO (bf16, d=128, causal).  Prints ONE JSON line (rank 0).  `--workload auto` (the default) selects:

  N = 1   value = the north-star headline shape B4 H16 N8192 (BASELINE "bf16 d=128 N=8K causal"), K timed steps
          (burst window) + a `sustained` block (>= 1 s of the same steps, clocks sampled) + a `c2` block (BASELINE
          configs[1], B4 H16 N4096) + a `c4` block (configs[3] whole on one GPU: the strong-scaling base).
  N > 1   value = BASELINE configs[3]: B32 H32 N8192 TOTAL, batch*head slices split over the N ranks, no data-path
          collective ("scaling": "strong"), + a `ring_c5` block: configs[4], N=131072 causal ring attention over the
          same N ranks (K/V and dK/dV blocks exchanged between neighbours, partials merged by LSE) with its own
          timing, transport, NVLink bytes and a PARITY field (ring vs the single-GPU kernel per tensor, and ring vs an
          fp32 dense reference on sampled query rows against all 131072 keys).

  value      whole-job fwd+bwd TFLOP/s, algorithmic FLOPs 14*B*H*N^2*d*(1/2 causal), inputs resident in HBM,
             CUDA-event timed over exactly K steps, max over ranks.
  e2e        the same metric through the public API with HOST (pinned) buffers: one packed H2D copy of q/k/v/dO and
             one packed D2H copy of o/dq/dk/dv/lse inside the timed region of every step; also reports the achieved
             GB/s per direction and the copy-only ceiling of the same buffers (no kernels).
  roofline   dominant kernel (the backward main kernel): algorithmic FLOPs / CUDA-event duration vs the measured
             cuBLAS bf16 peak of MEASURED_PEAKS.json (burst for the K-step window; the `sustained` block compares with
             the sustained peak).  `traffic` comes from the ncu --set full capture of this round (profiles/traffic.json).
  cpu_baseline  the oracle's CPU restatement ("port") timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's own native extension compiled for CPU (oracle/_ref, built from
/root/reference/csrc by oracle/build_ref.py; falls back to the oracle port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]

import torch  # noqa: E402

WORKLOADS = {
    # name: (B, H, N, d, causal)
    "c2": (4, 16, 4096, 128, True),        # BASELINE configs[1]
    "headline": (4, 16, 8192, 128, True),  # north-star "bf16 d=128 N=8K causal"
    "c4": (32, 32, 8192, 128, True),       # BASELINE configs[3]: TOTAL shape, batch*head split over the ranks
    "c5": (1, 16, 131072, 128, True),      # BASELINE configs[4]: TOTAL shape, sequence split over the ranks (ring)
}
NOMINAL_BF16_TFLOPS = 2250.0
FALLBACK_BF16_TFLOPS = 1590.0
OUR_KERNELS_PER_STEP = 4  # fa_fwd_kernel, fa_bwd_prepare_kernel (also zero-fills dq_accum), fa_bwd_kernel, fa_dq_finish_kernel


def flops(b, h, n, d, causal):
    c = 0.5 if causal else 1.0
    f_fwd = 4.0 * b * h * n * n * d * c
    return f_fwd, 2.5 * f_fwd


def default_workload(world):
    return "headline" if world == 1 else "c4"


def config_for(name, world):
    """The `config` object: identical in the GPU arm and the reference arm."""
    b, h, n, d, causal = WORKLOADS[name]
    return {"workload": name, "B": b, "H": h, "N": n, "d": d, "causal": causal, "shards": world}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"burst": float(j["bf16_tflops"]), "sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "source": "measured"}
    return {"burst": FALLBACK_BF16_TFLOPS, "sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------------
_NVML_SAMPLER_SRC = r"""
import sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
smax = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
out = open(sys.argv[2], "w", buffering=1)
out.write("ready %f\n" % time.time())
while True:
    try:
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        mask = int(reasons(h))
        out.write("%f %d %d %.3f %d\n" % (time.time(), sm, smax, pw, mask))
    except Exception:
        pass
    time.sleep(0.001)
"""


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region by a separate process polling NVML every
    ~1-3 ms (its own interpreter: the bench's launch loop cannot starve it), so even a 20 ms region gets several
    samples; `nvidia-smi -lms` is the fallback when the NVML bindings are unusable."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread, self.path, self.mode = index, [], None, None, None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip() != ""]
        if ids and self.index < len(ids) and ids[self.index].strip().isdigit():
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml  # noqa: F401  (only to know the child can import it)
            import tempfile
            fd, self.path = tempfile.mkstemp(prefix="fa_clocks_", suffix=".txt")
            os.close(fd)
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_SAMPLER_SRC, str(self._physical_index()), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t_end = time.time() + 5.0
            while time.time() < t_end:  # wait until the child has initialised NVML and started polling
                if os.path.getsize(self.path) > 0:
                    self.mode = "nvml"
                    return
                if self.proc.poll() is not None:
                    break
                time.sleep(0.01)
            self.proc.kill()
        except Exception:  # noqa: BLE001
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.mode = "nvidia-smi"
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        if self.mode == "nvml":
            time.sleep(0.01)
            self.proc.kill()
            self.proc.wait()
            pad = 0.0
            try:
                for line in open(self.path):
                    f = line.split()
                    if len(f) == 5:
                        mask = int(f[4])
                        flags = ["Active" if mask & self.BITS[k] else "Not Active" for k in self.NAMES]
                        self.rows.append((float(f[0]), [f[1], f[2], f[3], *flags]))
                os.unlink(self.path)
            except OSError:
                pass
        else:
            time.sleep(0.15)
            self.proc.terminate()
            pad = 0.2
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + pad]
        rows = inside or [r for _, r in self.rows]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        reasons = sorted({n for r in rows for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows),
                "samples_inside_timed_region": len(inside), "source": self.mode, "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------------------------------------------------
def _load_ref_ext():
    so = sorted((ROOT / "oracle" / "_ref").glob("flashattention_lab_cuda_ref*.so"))
    if not so:
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location("flashattention_lab_cuda_ref", so[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:  # noqa: BLE001  (ABI mismatch on another image -> fall back to the port)
        return None
    return mod


def cpu_step_fn(kind, n, d, causal, bh_sample):
    """Returns (callable running one fwd+bwd on the CPU over `bh_sample` slices, kind actually used)."""
    g = torch.Generator().manual_seed(0)
    q, k, v, do = (torch.randn((bh_sample, n, d), generator=g, dtype=torch.float32) for _ in range(4))
    scale = d ** -0.5
    ref = _load_ref_ext() if kind == "reference" else None
    if ref is not None:
        br, bc = (128, 128) if d <= 64 else (64, 128)  # reference src/fa2/spec.py

        def step():
            o, lse = ref.forward(q, k, v, causal, scale, br, bc)  # FA2 names, reference csrc/common/torch.extension.cpp:78-79
            return ref.backward(q, k, v, o, do, lse, causal, scale, br, bc)

        return step, "reference"
    from oracle.attention_oracle import blocked_backward, blocked_forward

    def step():
        o, lse = blocked_forward(q, k, v, causal, scale)
        return blocked_backward(q, k, v, o, do, lse, causal, scale)

    return step, "port"


def time_cpu(kind, workload, steps, warmup, budget_s=25.0):
    b, h, n, d, causal = workload
    bh_sample = 1
    step, used = cpu_step_fn(kind, n, d, causal, bh_sample)
    f_fwd, f_bwd = flops(1, bh_sample, n, d, causal)
    t0 = time.perf_counter()
    step()  # first call doubles as a cost estimate
    first = time.perf_counter() - t0
    steps = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    warmup = max(0, min(warmup, int(budget_s / 4 / max(first, 1e-3))))
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return {"value": (f_fwd + f_bwd) / mean / 1e12, "unit": "TFLOP/s", "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "kind": used, "ms_per_step_sample": mean * 1e3, "steps": steps,
            "sample": f"1 of {b * h} batch*head slices of the workload (N={n}, d={d}, causal={causal}, fp32 on CPU); "
                      f"throughput is per-slice work / time, slices are independent"}


def time_cpu_c1(kind):
    """BASELINE configs[0] exactly: B2 H4 N512 d64 fp32, causal and non-causal, fwd+bwd on the host cores."""
    out = {}
    for causal in (False, True):
        step, used = cpu_step_fn(kind, 512, 64, causal, 8)
        step()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
        f_fwd, f_bwd = flops(2, 4, 512, 64, causal)
        out["causal" if causal else "non_causal"] = {"min_ms": min(ts) * 1e3, "mean_ms": sum(ts) / len(ts) * 1e3,
                                                     "tflops": (f_fwd + f_bwd) / min(ts) / 1e12, "kind": used}
    return out


def run_reference_arm(args, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    workload = WORKLOADS[name]
    res = time_cpu("reference", workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "attention fwd+bwd TFLOP/s", "value": res["value"], "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": res["steps"], "warmup": args.warmup,
        "ms_per_step": res["ms_per_step_sample"], "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_for(name, world),
        "note": "reference csrc (ATen loops) compiled for CPU; its causal backward skips the wrong triangle "
                "(SURVEY D4) - same tile count, so the timing stands",
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "c1_B2_H4_N512_d64_fp32": time_cpu_c1("reference"),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class Env:
    """Rank / device / process-group plumbing of one bench process."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.host = pin_to_gpu_cpus(self.local)  # before any pinned allocation: first touch decides the NUMA node
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            # NCCL's copy kernels compete with the attention CTAs for SM slots: put them on a high-priority stream
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev, pg_options=opts)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def pin_to_gpu_cpus(local: int):
    """Bind this process to the CPUs NVML reports as local to its GPU, so pinned staging buffers are allocated on the
    GPU's NUMA node (round 1: every rank ran wherever the launcher put it)."""
    info = {"cpu_affinity": "unchanged"}
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
        phys = int(vis[local]) if local < len(vis) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < (os.cpu_count() or 1)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"cpu_affinity": f"{cpus[0]}-{cpus[-1]}" if cpus == list(range(cpus[0], cpus[-1] + 1)) else str(cpus),
                    "n_cpus": len(cpus)}
    except Exception as exc:  # noqa: BLE001  (no NVML / restricted container: keep the launcher's placement)
        info["note"] = repr(exc)[:80]
    return info


def timed_steps(env: Env, run_step, steps: int, sample: bool = True):
    """Exactly `steps` calls between barrier+synchronize, CUDA events on the launching stream, max over ranks; the SM
    clock / power / throttle reasons are sampled on rank 0 during the region."""
    sampler = ClockSampler(env.local)
    env.barrier()
    if sample and env.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        run_step()
    e1.record()
    env.barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if (sample and env.rank == 0) else None
    return env.max_over_ranks(e0.elapsed_time(e1)) / steps, clocks


def capture_graph(step, enabled: bool):
    """The step is 4 short launches (~1 ms of GPU work at C2); replaying it from a CUDA graph keeps the timed loop from
    depending on the host (8 ranks share one box's cores).  Same public-API call path, recorded then replayed."""
    if not enabled:
        return step, "eager"
    try:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            step()  # allocator warm-up on the side stream
            with torch.cuda.graph(graph, stream=cap):
                step()
        torch.cuda.current_stream().wait_stream(cap)
        graph.replay()
        torch.cuda.synchronize()
        return graph.replay, "cuda_graph"
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] CUDA graph capture failed ({exc!r}); timing eager launches", file=sys.stderr)
        torch.cuda.synchronize()
        return step, "eager"


def kernel_breakdown(ext, q, k, v, do, causal, scale, f_fwd, f_bwd, budget_ms=400.0):
    """Per-kernel CUDA-event times on the launching stream (raw entry points, same launches the public API makes)."""
    bh = q.shape[0] * q.shape[1]
    n, d = q.shape[2], q.shape[3]
    qb, kb, vb, dob = (x.detach().reshape(bh, n, d) for x in (q, k, v, do))
    ob, lseb = ext.fwd_raw(qb, kb, vb, causal, scale)
    dq_acc = torch.empty((bh, n, d), device=q.device, dtype=torch.float32)
    rowstats = ext.bwd_prepare_raw(ob, dob, lseb, zero=dq_acc)

    def time_kernel(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        iters = int(max(3, min(20, budget_ms / max(a.elapsed_time(b), 1e-3))))
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    t_fwd = time_kernel(lambda: ext.fwd_raw(qb, kb, vb, causal, scale, out=ob, lse=lseb))
    t_prep = time_kernel(lambda: ext.bwd_prepare_raw(ob, dob, lseb, zero=dq_acc))
    t_bwd = time_kernel(lambda: ext.bwd_raw(qb, kb, vb, ob, dob, lseb, causal, scale, rowstats=rowstats, dq_accum=dq_acc))
    t_fin = time_kernel(lambda: ext.dq_finish_raw(dq_acc, torch.bfloat16, scale))
    return {
        "fa_fwd_kernel": {"ms": t_fwd, "tflops": f_fwd / (t_fwd * 1e-3) / 1e12},
        "fa_bwd_prepare_kernel": {"ms": t_prep, "note": "delta + row statistics + zero-fill of the fp32 dQ accumulator"},
        "fa_bwd_kernel": {"ms": t_bwd, "tflops": f_bwd / (t_bwd * 1e-3) / 1e12},
        "fa_dq_finish_kernel": {"ms": t_fin},
        "bwd_total_tflops": f_bwd / ((t_prep + t_bwd + t_fin) * 1e-3) / 1e12,
    }


def measure_e2e(env: Env, fa2_attention, shape, causal, scale, f_step, steps):
    """fwd+bwd through the public API with HOST buffers: per step ONE packed H2D copy (q|k|v|dO) from pinned memory and
    ONE packed D2H copy (o|dq|dk|dv|lse) back, on three streams so step s+1's upload and step s-1's download overlap
    step s's kernels (PCIe is full duplex); every byte of every step crosses the bus inside the timed region."""
    dev = env.dev
    numel = math.prod(shape)
    t_bytes = numel * 2
    lse_bytes = math.prod(shape[:-1]) * 4
    h_in = torch.empty(4 * t_bytes, dtype=torch.uint8).pin_memory()
    h_in.view(torch.bfloat16).normal_()
    h_out = torch.empty(4 * t_bytes + lse_bytes, dtype=torch.uint8).pin_memory()
    d_in = [torch.empty(4 * t_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_out = [torch.empty(4 * t_bytes + lse_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]

    def views(buf):
        return [buf[i * t_bytes:(i + 1) * t_bytes].view(torch.bfloat16).view(shape) for i in range(4)]

    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_ready, in_free, packed, out_done = ([torch.cuda.Event() for _ in range(2)] for _ in range(4))

    def run(nsteps, with_kernels=True):
        for e in in_free + out_done:
            e.record(main)
        for s_ in range(nsteps):
            slot = s_ & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[slot])      # the kernels that last read this device buffer are done
                d_in[slot].copy_(h_in, non_blocking=True)
                in_ready[slot].record(s_in)
            main.wait_event(in_ready[slot])
            main.wait_event(out_done[slot])         # the previous download out of d_out[slot] is done
            if with_kernels:
                q_, k_, v_, do_ = views(d_in[slot])
                q_, k_, v_ = (x.detach().requires_grad_(True) for x in (q_, k_, v_))
                o, lse = fa2_attention(q_, k_, v_, causal=causal, softmax_scale=scale, backend="cuda")
                torch.autograd.backward(o, do_)
                for dst, src in zip(views(d_out[slot]), (o.detach(), q_.grad, k_.grad, v_.grad)):
                    dst.copy_(src)                  # pack on the device: ~0.1 ms per 100 MB, then ONE download
                d_out[slot][4 * t_bytes:].view(torch.float32).view(shape[:-1]).copy_(lse.detach())
            in_free[slot].record(main)
            packed[slot].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(packed[slot])
                h_out.copy_(d_out[slot], non_blocking=True)
                out_done[slot].record(s_out)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    def timed(nsteps, with_kernels):
        run(2, with_kernels)
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(nsteps, with_kernels)
        e1.record()
        env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1)) / nsteps

    ms = timed(steps, True)
    ms_copy = timed(steps, False)
    h2d, d2h = int(h_in.numel()), int(h_out.numel())
    return {"value": f_step * env.world / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "h2d_gbps_per_gpu": h2d / (ms * 1e-3) / 1e9, "d2h_gbps_per_gpu": d2h / (ms * 1e-3) / 1e9,
            "copy_only": {"ms_per_step": ms_copy, "h2d_gbps_per_gpu": h2d / (ms_copy * 1e-3) / 1e9,
                          "d2h_gbps_per_gpu": d2h / (ms_copy * 1e-3) / 1e9,
                          "note": "the same packed uploads/downloads with no kernels: the host-link ceiling of this box "
                                  "at this rank count"},
            "host": env.host,
            "note": "pinned host buffers, 1 packed H2D + 1 packed D2H per step; uploads, kernels and downloads of "
                    "consecutive steps overlap on three streams"}


def measure_bh_workload(env: Env, args, name, *, split_over_ranks, steps, e2e_steps, sustain_s, with_kernels):
    """fwd+bwd of `fa2_attention` over this rank's batch*head slices of workload `name`."""
    import flashattention_lab_cuda as ext
    from dist.shard import shard_range
    from fa2 import fa2_attention

    b, h, n, d, causal = WORKLOADS[name]
    if split_over_ranks:
        lo, hi = shard_range(b * h, env.rank, env.world)
        b_local, h_local = 1, hi - lo
    else:
        b_local, h_local = b, h
    f_fwd, f_bwd = flops(b_local, h_local, n, d, causal)
    f_total = sum(flops(b, h, n, d, causal)) if split_over_ranks else (f_fwd + f_bwd) * env.world
    scale = d ** -0.5
    shape = (b_local, h_local, n, d)
    g = torch.Generator(device=env.dev).manual_seed(env.rank)
    q, k, v = (torch.randn(shape, generator=g, device=env.dev, dtype=torch.bfloat16).requires_grad_(True)
               for _ in range(3))
    do = torch.randn(shape, generator=g, device=env.dev, dtype=torch.bfloat16)

    def step():
        o, _ = fa2_attention(q, k, v, causal=causal, softmax_scale=scale, backend="cuda")
        torch.autograd.backward(o, do)
        q.grad = k.grad = v.grad = None

    for _ in range(max(args.warmup, 3)):
        step()
    out = {"B_per_gpu": b_local, "H_per_gpu": h_local}
    if with_kernels:
        out["kernels"] = kernel_breakdown(ext, q, k, v, do, causal, scale, f_fwd, f_bwd)
    if e2e_steps:
        out["e2e"] = measure_e2e(env, fa2_attention, shape, causal, scale, f_fwd + f_bwd, e2e_steps)
    run_step, out["launch"] = capture_graph(step, not args.no_graph)
    # the K-step (burst) window is measured after everything else of this workload, so nothing perturbs it
    ms_step, clocks = timed_steps(env, run_step, steps)
    if env.world == 1 and (clocks is None or not clocks.get("samples_inside_timed_region")):
        # the region was too short for the sampler (or it started late): sample over >= 0.2 s of the same steps
        extra = ClockSampler(env.local)
        extra.start()
        tx0 = time.time()
        while time.time() - tx0 < 0.2:
            for _ in range(steps):
                run_step()
            torch.cuda.synchronize()
        again = extra.stop(tx0, time.time())
        if again is not None:
            again["note"] = "sampled over extra replays of the same steps right after the timed region"
            clocks = again
    out.update({"value": f_total / (ms_step * 1e-3) / 1e12, "ms_per_step": ms_step, "steps": steps, "clocks": clocks,
                "flops_per_step_total": f_total})
    if sustain_s > 0:
        n_sus = max(steps, int(math.ceil(sustain_s * 1e3 / ms_step)))
        ms_sus, clocks_sus = timed_steps(env, run_step, n_sus)
        peaks = measured_peaks()
        v_sus = f_total / (ms_sus * 1e-3) / 1e12
        out["sustained"] = {"value": v_sus, "unit": "TFLOP/s", "steps": n_sus, "ms_per_step": ms_sus,
                            "seconds": n_sus * ms_sus * 1e-3, "clocks": clocks_sus,
                            "frac_of_measured_sustained_peak": v_sus / env.world / peaks["sustained"],
                            "frac_of_nominal_bf16_peak": v_sus / env.world / NOMINAL_BF16_TFLOPS}
    del q, k, v, do
    torch.cuda.empty_cache()
    return out


def roofline_of(kernels, name, f_bwd_per_launch):
    peaks = measured_peaks()
    ach = kernels["fa_bwd_kernel"]["tflops"]
    traffic, src = None, None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            j = json.loads(tf.read_text())
            traffic = j.get(name, {}).get("fa_bwd_kernel_dram_bytes_per_launch")
            src = j.get(name, {}).get("source")
        except Exception:  # noqa: BLE001
            traffic = None
    return {"kernel": "fa_bwd_kernel<128,bf16>", "bound": "tensor", "achieved": ach, "peak": peaks["burst"],
            "unit": "TFLOP/s", "frac": ach / peaks["burst"], "peak_source": peaks["source"] + " (burst: kernel timed alone)",
            "frac_of_sustained": ach / peaks["sustained"], "frac_of_nominal": ach / NOMINAL_BF16_TFLOPS,
            "algorithmic_flops_per_launch": f_bwd_per_launch, "launch_ms": kernels["fa_bwd_kernel"]["ms"],
            "traffic": traffic, "traffic_source": src}


def ring_block(env: Env, args):
    """BASELINE configs[4]: N=131072 causal ring attention over the same ranks, with timing and parity."""
    import flashattention_lab_cuda as ext
    from dist.ring import ring_attention, ring_transport_name, zigzag_split

    b, h, n, d, causal = WORKLOADS["c5"]
    world, dev = env.world, env.dev
    if n % (2 * world * 128):
        return {"skipped": "N must be a multiple of 256 * n_gpus"}
    n_local = n // world
    bh = b * h
    scale = d ** -0.5
    f_fwd, f_bwd = flops(b, h, n, d, causal)
    # the same GLOBAL tensors on every rank (same seed), each rank keeps its zig-zag rows
    g = torch.Generator(device=dev).manual_seed(1234)
    glob = [torch.randn((bh, n, d), generator=g, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    q, k, v = (zigzag_split(t, world)[env.rank].requires_grad_(True) for t in glob[:3])
    do = zigzag_split(glob[3], world)[env.rank]

    def step():
        o, lse = ring_attention(q, k, v, causal=causal, softmax_scale=scale)
        torch.autograd.backward(o, do)
        grads = (q.grad, k.grad, v.grad)
        q.grad = k.grad = v.grad = None
        return o, lse, grads

    for _ in range(2):
        step()
    steps = max(3, min(args.steps, 10))
    ms_step, clocks = timed_steps(env, step, steps)
    value = (f_fwd + f_bwd) / (ms_step * 1e-3) / 1e12

    # ---- parity (outside the timed region) ----
    o, lse, (dq, dk, dv) = step()
    torch.cuda.synchronize()
    # (a) against the single-GPU kernel on the full sequence (every rank computes it and checks its own rows)
    gq, gk, gv, gdo = glob
    o_ref, lse_ref = ext.fwd_raw(gq, gk, gv, causal, scale)
    dq_ref, dk_ref, dv_ref = ext.bwd_raw(gq, gk, gv, o_ref, gdo, lse_ref, causal, scale)
    mine = lambda t: zigzag_split(t, world)[env.rank]  # noqa: E731
    vs_kernel = {}
    for nm, got, ref in (("o", o, o_ref), ("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        vs_kernel[nm] = env.max_over_ranks((got.float() - mine(ref).float()).abs().max().item())
    vs_kernel["lse"] = env.max_over_ranks((lse - mine(lse_ref.unsqueeze(-1)).squeeze(-1)).abs().max().item())
    del o_ref, dq_ref, dk_ref, dv_ref
    # (b) against an fp32 dense reference on sampled query rows of THIS rank vs all keys (torch fp32 on the GPU)
    rows_per_rank = max(8, 256 // world)
    c = n_local // 2
    local_idx = torch.linspace(0, n_local - 1, rows_per_rank, device=dev).long().unique()
    chunk_a, chunk_b = env.rank, 2 * world - 1 - env.rank
    global_idx = torch.where(local_idx < c, chunk_a * c + local_idx, chunk_b * c + (local_idx - c))
    worst_o = worst_lse = 0.0
    kf_all, vf_all = gk.float(), gv.float()
    for s0 in range(0, bh, 4):  # 4 heads at a time: [4, rows, 131072] fp32 scores
        qs = q.detach()[s0:s0 + 4, local_idx].float()
        sc = torch.einsum("hrd,hnd->hrn", qs, kf_all[s0:s0 + 4]) * scale
        cols = torch.arange(n, device=dev)
        sc.masked_fill_(cols[None, None, :] > global_idx[None, :, None], float("-inf"))
        lse_d = torch.logsumexp(sc, dim=-1)
        o_d = torch.einsum("hrn,hnd->hrd", torch.softmax(sc, dim=-1), vf_all[s0:s0 + 4])
        worst_o = max(worst_o, (o.detach()[s0:s0 + 4, local_idx].float() - o_d).abs().max().item())
        worst_lse = max(worst_lse, (lse.detach()[s0:s0 + 4, local_idx] - lse_d).abs().max().item())
        del sc, o_d
    vs_dense = {"rows": int(rows_per_rank * world), "keys": n, "o_max_abs": env.max_over_ranks(worst_o),
                "lse_max_abs": env.max_over_ranks(worst_lse), "tolerance": {"o": 5e-2, "lse": 1e-3}}
    ok = (max(vs_kernel.values()) < 5e-2 and vs_dense["o_max_abs"] < 5e-2 and vs_dense["lse_max_abs"] < 1e-3)
    kv_bytes = 2 * bh * n_local * d * 2
    return {"workload": "c5", "metric": "attention fwd+bwd TFLOP/s", "value": value, "unit": "TFLOP/s",
            "per_gpu_tflops": value / world, "ms_per_step": ms_step, "steps": steps, "scaling": "strong",
            "config": config_for("c5", world), "N_per_gpu": n_local, "transport": ring_transport_name(q),
            "nvlink_bytes_per_gpu_per_step": (world - 1) * kv_bytes + world * (kv_bytes + 2 * kv_bytes),
            "frac_of_nominal_bf16_peak": value / world / NOMINAL_BF16_TFLOPS, "clocks": clocks,
            "parity": {"ok": bool(ok), "ring_vs_single_gpu_kernel_max_abs": vs_kernel,
                       "ring_vs_fp32_dense_rows": vs_dense}}


def run_ours(args, name):
    env = Env()
    import flashattention_lab_cuda as ext

    ext.load_library()  # fail loudly if the CUDA library is missing
    auto = args.workload == "auto"
    world = env.world
    b, h, n, d, causal = WORKLOADS[name]
    split = name == "c4"
    scaling = "strong" if (split and world > 1) else "weak"
    if split and (b * h) % world:
        raise SystemExit("c4: batch*head slices must divide evenly over the ranks")
    e2e_steps = max(3, min(args.steps, 10 if world == 1 else 3))
    main_res = measure_bh_workload(env, args, name, split_over_ranks=split, steps=args.steps, e2e_steps=e2e_steps,
                                   sustain_s=args.sustain_seconds if world == 1 else 0.0, with_kernels=True)
    extras = {}
    if auto and world == 1:
        quick = dict(e2e_steps=0, sustain_s=0.0)
        c2 = measure_bh_workload(env, args, "c2", split_over_ranks=False, steps=args.steps, with_kernels=True, **quick)
        f2 = flops(*WORKLOADS["c2"])
        extras["c2"] = {"config": config_for("c2", 1), "value": c2["value"], "unit": "TFLOP/s",
                        "ms_per_step": c2["ms_per_step"], "steps": c2["steps"], "clocks": c2["clocks"],
                        "kernels": c2["kernels"], "roofline": roofline_of(c2["kernels"], "c2", f2[1]),
                        "frac_of_nominal_bf16_peak": c2["value"] / NOMINAL_BF16_TFLOPS}
        c4 = measure_bh_workload(env, args, "c4", split_over_ranks=True, steps=max(3, min(args.steps, 5)),
                                 with_kernels=False, **quick)
        extras["c4"] = {"config": config_for("c4", 1), "value": c4["value"], "unit": "TFLOP/s",
                        "ms_per_step": c4["ms_per_step"], "steps": c4["steps"], "clocks": c4["clocks"],
                        "note": "BASELINE configs[3] whole on one GPU: the base of the strong-scaling curve that "
                                "`bench.py --gpus N` (N > 1) reports as `value`"}
    if auto and world > 1 and not args.no_ring:
        try:
            extras["ring_c5"] = ring_block(env, args)
        except Exception as exc:  # noqa: BLE001  (keep the C4 line even if the ring leg fails; say so loudly)
            extras["ring_c5"] = {"error": repr(exc)[:400]}
            print(f"[bench] ring_c5 failed on rank {env.rank}: {exc!r}", file=sys.stderr)

    if env.rank == 0:
        peaks = measured_peaks()
        f_bwd_launch = flops(main_res["B_per_gpu"], main_res["H_per_gpu"], n, d, causal)[1]
        line = {
            "metric": "attention fwd+bwd TFLOP/s", "value": main_res["value"], "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_for(name, world),
            "config_detail": {"B_per_gpu": main_res["B_per_gpu"], "H_per_gpu": main_res["H_per_gpu"],
                              "parallelism": f"batch*head slices over {world} rank(s), no data-path collective",
                              "launch": main_res["launch"], "flop_convention": "14*B*H*N^2*d*(0.5 if causal)",
                              "l2": "inputs (q,k,v,dO >= 4 x 64 MiB) exceed the 126 MB L2; no explicit flush"},
            "frac_of_nominal_bf16_peak": main_res["value"] / world / NOMINAL_BF16_TFLOPS,
            "frac_of_measured_bf16_peak": main_res["value"] / world / peaks["burst"],
            "clocks": main_res["clocks"], "e2e": main_res["e2e"],
            "gpu_launches": OUR_KERNELS_PER_STEP * args.steps,
            "launches_per_step": {"ours": OUR_KERNELS_PER_STEP, "other": 0,
                                  "names": ["fa_fwd_kernel", "fa_bwd_prepare_kernel", "fa_bwd_kernel", "fa_dq_finish_kernel"]},
            "roofline": roofline_of(main_res["kernels"], name, f_bwd_launch), "kernels": main_res["kernels"],
        }
        if "sustained" in main_res:
            line["sustained"] = main_res["sustained"]
        line.update(extras)
        if world == 1:
            cpu = time_cpu("port", WORKLOADS[name], 3, 1)
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    env.close()


def run_ring(args, name):
    """`--workload c5`: only the ring leg (BASELINE configs[4]) as the line's value."""
    env = Env()
    if env.world < 2:
        raise SystemExit("c5 needs --gpus >= 2 (ring attention shards the sequence over the ranks)")
    blk = ring_block(env, args)
    if env.rank == 0:
        line = {"metric": blk["metric"], "value": blk["value"], "unit": "TFLOP/s", "n_gpus": env.world,
                "steps": blk["steps"], "warmup": 2, "ms_per_step": blk["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": blk["config"],
                "config_detail": {k: blk[k] for k in ("N_per_gpu", "transport", "nvlink_bytes_per_gpu_per_step")},
                "frac_of_nominal_bf16_peak": blk["frac_of_nominal_bf16_peak"], "clocks": blk["clocks"],
                "parity": blk["parity"], "gpu_launches": None}
        print(json.dumps(line), flush=True)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    ap.add_argument("--no-ring", action="store_true", help="N > 1, auto: skip the ring_c5 block")
    ap.add_argument("--sustain-seconds", type=float, default=1.2,
                    help="N = 1: length of the extra sustained-throughput region (0 disables it)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = default_workload(max(world, args.gpus)) if args.workload == "auto" else args.workload
    if args.impl == "reference":
        run_reference_arm(args, name)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback (use --impl reference)")
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--workload", args.workload, "--sustain-seconds", str(args.sustain_seconds)]
        cmd += (["--no-graph"] if args.no_graph else []) + (["--no-ring"] if args.no_ring else [])
        raise SystemExit(subprocess.call(cmd))
    if name == "c5":
        run_ring(args, name)
        return
    run_ours(args, name)


if __name__ == "__main__":
    main()
