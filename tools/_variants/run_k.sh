timeout 900 python -m pytest tests/test_kernel_parity.py tests/test_ring_gpu.py -m gpu -x -q 2>&1 | tail -8
