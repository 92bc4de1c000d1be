timeout 900 python -m pytest tests/test_block_sparse_dropout.py -m gpu -x -q 2>&1 | tail -30
timeout 900 python -m pytest tests/test_kernel_parity.py tests/test_correctness_fa2.py tests/test_ring_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/quick_perf.py --no-parity --short
