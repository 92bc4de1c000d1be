#!/bin/bash
# quick A/B under gpurun: parity subset + kernel rates for the in-tree library and any tools/_variants/lib_*.so given
# usage: tools/gpu_quick.sh <tag> [variant names...]   (env QUICK_TESTS=1 also runs the kernel parity test files)
set -u
tag=$1; shift
mkdir -p gpurun_out
{
echo "== in-tree =="; timeout 300 python tools/quick_perf.py
for v in "$@"; do echo "== variant $v =="; FA_SM100_LIB=tools/_variants/lib_$v.so timeout 300 python tools/quick_perf.py --no-parity; done
if [ "${QUICK_TESTS:-0}" = "1" ]; then timeout 900 python -m pytest tests/test_kernel_parity.py tests/test_ring_gpu.py tests/test_correctness_fa2.py -m gpu -x -q 2>&1 | tail -15; fi
} 2>&1 | tee gpurun_out/${tag}_quick.log
