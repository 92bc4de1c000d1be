"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the keys
the driver reads; the clock sampler parses what its NVML child process writes."""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "attention fwd+bwd TFLOP/s" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"] == {"workload": "headline", "B": 4, "H": 16, "N": 8192, "d": 128, "causal": True, "shards": 1}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_clock_sampler_parses_child_output(tmp_path):
    sys.path.insert(0, str(ROOT))
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    t0 = time.time()
    path = tmp_path / "clocks.txt"
    # ready line, two samples inside the region (the second one power-capped), one after it
    path.write_text(f"ready {t0}\n{t0 + 0.01} 1950 1965 300.5 0\n{t0 + 0.02} 1900 1965 310.0 4\n{t0 + 9.0} 1000 1965 90.0 0\n")

    class _Done:  # stands in for the finished child process
        def kill(self): pass
        def wait(self): pass

    s.proc, s.path, s.mode = _Done(), str(path), "nvml"
    out = s.stop(t0, t0 + 1.0)
    assert out["samples_inside_timed_region"] == 2 and out["sm_mhz"] == 1925.0 and out["sm_max_mhz"] == 1965.0
    assert out["power_w_max"] == 310.0 and out["reasons"] == ["sw_power_cap"] and out["source"] == "nvml"


def test_sweep_records_follow_the_reference_benchmark_schema(tmp_path):
    """tools/sweep.py writes records the reference's plotting can load: the BenchmarkRecord fields of reference
    benchmarks/bench_utils.py:161-180 in the CSV column order of write_results (:300-319), its FLOP convention
    (:210-215) and its CLI flags (:250-264)."""
    import argparse
    import csv
    import importlib.util

    spec = importlib.util.spec_from_file_location("sweep_under_test", ROOT / "tools" / "sweep.py")
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    reference_fields = ["method", "algo", "backend", "direction", "dtype", "causal", "seqlen", "head_dim", "batch_size",
                        "num_heads", "mean_ms", "std_ms", "tflops", "peak_mem_mb", "status", "fp8", "config", "error"]
    assert sweep.REF_FIELDS == reference_fields
    fwd = sweep.make_record("fa2", "forward", "bf16", True, 4096, 128, 4, 16, 2.0, 0.1, 123.0, algorithmic_tflops=500.0)
    both = sweep.make_record("fa3", "backward", "fp16", False, 1024, 64, 2, 4, 1.0, 0.0, 1.0)
    bad = sweep.make_record("fa1", "forward", "fp32", False, 512, 64, 1, 4, None, None, None, "unsupported", "dtype")
    for rec in (fwd, both, bad):
        assert list(rec) == reference_fields
    assert abs(fwd["tflops"] - 4.0 * 4 * 16 * 4096 ** 2 * 128 / 2e-3 / 1e12) < 1e-9  # no causal discount (reference)
    assert abs(both["tflops"] - 8.0 * 2 * 4 * 1024 ** 2 * 64 / 1e-3 / 1e12) < 1e-9 and both["fp8"] is False
    assert fwd["fp8"] is None and bad["tflops"] is None and bad["status"] == "unsupported"
    f8 = sweep.make_record("fa3", "forward", "bf16", True, 4096, 128, 4, 16, 2.0, 0.1, 1.0, fp8=True)
    assert f8["fp8"] is True and f8["method"].endswith(" FP8") and list(f8) == reference_fields  # compare_all's label
    sweep.write_results("t", [fwd, both, bad], tmp_path)
    rows = list(csv.DictReader((tmp_path / "t.csv").open()))
    assert list(rows[0]) == reference_fields and len(rows) == 3
    assert json.loads((tmp_path / "t.json").read_text())[0]["seqlen"] == 4096
    ap = argparse.ArgumentParser()
    sweep.add_common_args(ap)
    ns = ap.parse_args([])
    assert ns.seqlen == [512, 1024, 2048, 4096, 8192, 16384] and ns.head_dim == [64, 128, 256]
    assert ns.batch_size == [1, 2] and ns.num_heads == [4] and ns.dtypes == ["fp16", "bf16"]
    assert (ns.warmup, ns.iters) == (5, 20) and sweep.iter_causal_flags(ns) == [False, True]
