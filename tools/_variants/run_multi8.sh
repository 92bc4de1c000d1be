#!/bin/bash
set -u
n=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
python -m pytest tests/test_ring_multigpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${1}_ring_multigpu_${n}gpu.log
python bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${1}_bench_${n}gpu.json 2> gpurun_out/${1}_bench_${n}gpu.err
tail -c 1500 gpurun_out/${1}_bench_${n}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${1}_bench_${n}gpu.json") if l.startswith("{")][-1])
    print("C4", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["e2e"]["h2d_gbps_per_gpu"], d["e2e"]["copy_only"]["h2d_gbps_per_gpu"])
    r = d.get("ring_c5", {})
    print("RING", r.get("value"), r.get("ms_per_step"), r.get("transport"), r.get("parity"), r.get("clocks"), r.get("error"))
except Exception as e:
    print("no line", e)
PY
FA_RING_TRANSPORT=nccl python bench.py --gpus $n --workload c5 --steps 8 --warmup 2 > gpurun_out/${1}_c5_${n}gpu_nccl.json 2> gpurun_out/${1}_c5_${n}gpu_nccl.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${1}_c5_${n}gpu_nccl.json") if l.startswith("{")][-1])
    print("nccl", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), d["config_detail"]["transport"], d["clocks"], d["parity"]["ok"])
except Exception as e:
    print("nccl no line", e)
PY
