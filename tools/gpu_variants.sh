#!/bin/bash
# usage: gpu_variants.sh <kind: fwdperf|all> variant...   -> runs a parity smoke + perf for each variant library
kind=$1; shift
for v in "$@"; do
  export FA_SM100_LIB=$PWD/tools/_variants/lib_$v.so
  echo "=== variant $v"
  python tools/gpu_bringup.py fwd 2 1024 128 bfloat16 1
  python tools/gpu_bringup.py fwd 3 777 64 float16 0
  python tools/gpu_bringup.py perf 4 16 8192 128 1
  python tools/gpu_bringup.py perf 4 16 8192 128 0
  python tools/gpu_bringup.py perf 4 16 4096 128 1
done
