#!/bin/bash
# NCCL transport variants for the C5 ring at this GPU count (each a separate torchrun)
# every run under its own timeout: NCCL_P2P_USE_CUDA_MEMCPY=1 hangs in the first exchange at 8 GPUs (profiles/r02y_nccl_variants_8gpu.log)
set -u
n=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  env "$@" FA_RING_TRANSPORT=nccl timeout 150 python bench.py --gpus $n --workload c5 --steps 6 --warmup 2 > gpurun_out/r02y_c5_${n}gpu_$tag.json 2> gpurun_out/r02y_c5_${n}gpu_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02y_c5_${n}gpu_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), d["parity"]["ok"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag no line", e)
PY
  tail -2 gpurun_out/r02y_c5_${n}gpu_$tag.err | cut -c1-300
}
run default X=1
run cudamemcpy NCCL_P2P_USE_CUDA_MEMCPY=1
run maxctas2 NCCL_MAX_CTAS=2
run maxctas8 NCCL_MAX_CTAS=8 NCCL_MIN_CTAS=8
run nchan2 NCCL_MAX_NCHANNELS=2 NCCL_MIN_NCHANNELS=2
FA_RING_TRANSPORT=symm timeout 150 python bench.py --gpus $n --workload c5 --steps 6 --warmup 2 > gpurun_out/r02y_c5_${n}gpu_symm.json 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02y_c5_${n}gpu_symm.json") if l.startswith("{")][-1])
print("symm", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), d["parity"]["ok"], d["clocks"]["sm_mhz"])
PY
