#!/bin/bash
for v in "$@"; do
  export FA_SM100_LIB=$PWD/tools/_variants/lib_$v.so
  echo "=== variant $v"
  python tools/gpu_bringup.py perf 4 16 8192 128 0
  python tools/gpu_bringup.py perf 4 16 4096 128 1
done
