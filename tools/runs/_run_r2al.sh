timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "outer_loop" 2>&1 | tail -12 | cut -c1-300
