timeout 600 python -m pytest tests/test_fp8_path.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_kernel_parity.py -m gpu -q -k "head_dims_up_to_256" 2>&1 | tail -3
python tools/fp8_perf.py
bash tools/_variants/run_r.sh 2>&1 | tail -3
