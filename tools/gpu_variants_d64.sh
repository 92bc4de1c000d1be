#!/bin/bash
for v in "$@"; do
  export FA_SM100_LIB=$PWD/tools/_variants/lib_$v.so
  echo "=== variant $v"
  python tools/gpu_bringup.py fwd 3 777 64 float16 1
  python tools/gpu_bringup.py perf 4 32 4096 64 1
  python tools/gpu_bringup.py perf 4 32 8192 64 0
done
