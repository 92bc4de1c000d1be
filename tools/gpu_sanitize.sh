#!/bin/bash
# compute-sanitizer over the four small shapes of tools/sanitize_target.py: memcheck, racecheck, synccheck (one tool
# per process, bounded by `timeout`).  Logs -> gpurun_out/sanitizer_<tool>.log; copy summaries into profiles/.
set -u
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "target fails without sanitizer"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $? =="; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_OK|violations=" gpurun_out/sanitizer_$tool.log | tail -8
done
