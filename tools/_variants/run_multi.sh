#!/bin/bash
# multi-GPU session: real-rank ring parity test, then the default bench line at this GPU count (C4 strong + ring_c5)
set -u
n=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
python -m pytest tests/test_ring_multigpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${1}_ring_multigpu_${n}gpu.log
python bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${1}_bench_${n}gpu.json 2> gpurun_out/${1}_bench_${n}gpu.err
tail -c 2500 gpurun_out/${1}_bench_${n}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${1}_bench_${n}gpu.json") if l.startswith("{")][-1])
    print("C4", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["e2e"]["h2d_gbps_per_gpu"], d["e2e"]["copy_only"])
    print("RING", json.dumps(d.get("ring_c5"))[:1500])
except Exception as e:
    print("no line", e)
PY
