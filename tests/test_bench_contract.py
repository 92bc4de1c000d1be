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
    assert d["config"]["workload"] == "c2" and "model" not in d["config"]
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
