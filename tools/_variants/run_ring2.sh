#!/bin/bash
set -u
n=$(nvidia-smi -L | wc -l)
for tr in auto nccl; do
  FA_RING_TRANSPORT=$tr python bench.py --gpus $n --workload c5 --steps 8 --warmup 2 > gpurun_out/${1}_c5_${n}gpu_$tr.json 2> gpurun_out/${1}_c5_${n}gpu_$tr.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${1}_c5_${n}gpu_$tr.json") if l.startswith("{")][-1])
    print("$tr", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), d["config_detail"]["transport"], d["clocks"], d["parity"]["ok"])
except Exception as e:
    print("$tr no line", e)
PY
  tail -3 gpurun_out/${1}_c5_${n}gpu_$tr.err
done
