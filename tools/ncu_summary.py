"""Summarise an `ncu --set full` capture into profiles/<tag>_ncu_full_summary_bench_<workload>.json and refresh the
workload's entry of profiles/traffic.json (DRAM bytes per launch of the main kernels, read by bench.py's roofline).
usage: python tools/ncu_summary.py <rep> <tag> [workload=c2]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "prof.ncu-rep")
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
workload = sys.argv[3] if len(sys.argv) > 3 else "c2"
KEEP = ["sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
summary, traffic = {}, {}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


for r in data:
    name = r[ix["Kernel Name"]]
    key = name.split("(")[0]
    if key in summary:
        continue
    rec = {}
    for k in KEEP:
        if k in ix:
            rec[k] = f"{r[ix[k]]} {units[ix[k]]}".strip()
    fl = num(r[ix[KEEP[19]]]) if KEEP[19] in ix else None
    dur = num(r[ix["gpu__time_duration.sum"]])
    dur_unit = units[ix["gpu__time_duration.sum"]]
    dur_s = dur * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(dur_unit, 1e-6)
    clk = num(r[ix["sm__cycles_elapsed.max.per_second"]]) if "sm__cycles_elapsed.max.per_second" in ix else None
    clk_unit = units[ix["sm__cycles_elapsed.max.per_second"]] if clk else ""
    clk_hz = clk * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(clk_unit, 1e9) if clk else None
    if fl and dur_s:
        rec["derived_executed_tensor_tflops"] = fl / dur_s / 1e12
        if clk_hz:
            peak_at_clk = 2250e12 * clk_hz / 1.965e9  # nominal dense bf16 scales with the SM clock (max 1965 MHz)
            rec["derived_tensor_pipe_utilisation_at_measured_clock"] = fl / dur_s / peak_at_clk
            rec["derived_note"] = ("utilisation = executed UTCHMMA flops / (duration x 2250 TFLOP/s x measured SM clock "
                                   "/ 1965 MHz); ncu serialises and cold-starts kernels, so durations exceed the bench's")
    rd, wr = num(r[ix["dram__bytes_read.sum"]]), num(r[ix["dram__bytes_write.sum"]])
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    if rd is not None and wr is not None:
        traffic[key] = rd * mult.get(units[ix["dram__bytes_read.sum"]], 1.0) + wr * mult.get(units[ix["dram__bytes_write.sum"]], 1.0)
    summary[key] = rec
out_name = f"{tag}_ncu_full_summary_bench_{workload}.json"
(ROOT / "profiles" / out_name).write_text(json.dumps(summary, indent=1))
tpath = ROOT / "profiles" / "traffic.json"
tj = json.loads(tpath.read_text()) if tpath.exists() else {}
entry = {"source": f"profiles/{out_name} (ncu --set full --clock-control none on `python bench.py --workload {workload} "
                   f"--steps 2 --warmup 1 --sustain-seconds 0`, one capture, per launch)"}
for key, val in traffic.items():
    if "fa_bwd_kernel" in key:
        entry["fa_bwd_kernel_dram_bytes_per_launch"] = val
    if "fa_fwd_kernel" in key:
        entry["fa_fwd_kernel_dram_bytes_per_launch"] = val
tj[workload] = entry
tpath.write_text(json.dumps(tj, indent=1))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk.startswith("derived") or "duration" in kk or "dram__bytes" in kk}
                  for k, v in summary.items()}, indent=1))
print("traffic", traffic)
