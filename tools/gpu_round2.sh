#!/bin/bash
# Round-2 GPU session (1 GPU): tests, smoke, both bench arms, sanitizer, then the two ncu passes on the C2 step.
# usage: tools/gpu_round2.sh <tag> [tests] [bench] [sanitize] [ncu] [sweep]
set -u
tag=${1:-r02}; shift
mkdir -p gpurun_out
want() { for a in "$@"; do :; done; [[ " $ALL " == *" $1 "* ]]; }
ALL="$*"; [ -z "$ALL" ] && ALL="tests bench"
if want tests; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/${tag}_smoke.log
fi
if want bench; then
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err
  python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
k = d["kernels"]
print("HEADLINE", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "fwd", round(k["fa_fwd_kernel"]["tflops"]), "bwd", round(k["fa_bwd_kernel"]["tflops"]),
      "sustained", round(d.get("sustained", {}).get("value", 0)), "e2e", round(d["e2e"]["value"]), "c2", round(d.get("c2", {}).get("value", 0)),
      "c4", round(d.get("c4", {}).get("value", 0)), "clocks", d["clocks"])
PY
fi
if want sanitize; then bash tools/gpu_sanitize.sh; fi
if want sweep; then python tools/sweep.py --c3 --name ${tag}_sweep_records > gpurun_out/${tag}_sweep_c3.md 2> gpurun_out/${tag}_sweep.err; tail -5 gpurun_out/${tag}_sweep.err; grep -c "^|" gpurun_out/${tag}_sweep_c3.md; fi
if want ncu; then
  for wl in c2 headline; do
    B="python bench.py --workload $wl --steps 2 --warmup 1 --sustain-seconds 0"
    $B > gpurun_out/plain_$wl.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_$wl.csv $B > gpurun_out/ncu1_$wl.log 2>&1
    $B > gpurun_out/plain2_$wl.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:fa_ -s 8 -c 4 -f -o gpurun_out/${tag}_prof_$wl $B > gpurun_out/ncu2_$wl.log 2>&1
    tail -2 gpurun_out/ncu1_$wl.log gpurun_out/ncu2_$wl.log
  done
fi
if want parity; then python tools/parity_report.py > gpurun_out/${tag}_parity_report.md 2>&1; tail -3 gpurun_out/${tag}_parity_report.md; fi
if want sweep256; then python tools/sweep.py --algos fa2 --dtypes bf16 --head-dim 64 128 256 --seqlen 1024 4096 8192 --batch-size 4 --num-heads 16 --warmup 3 --iters 10 --name ${tag}_sweep_headdims > gpurun_out/${tag}_sweep_headdims.md 2> gpurun_out/${tag}_sweep_headdims.err; tail -3 gpurun_out/${tag}_sweep_headdims.err; grep -c "^|" gpurun_out/${tag}_sweep_headdims.md; fi
