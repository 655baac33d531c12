# dev tool: phase times of the loop kernel on three representative factors, the GPU parity tests and the bench
timeout 120 python tools/profile_target.py 300 layer4.1.conv1 36 1 0 2>&1 | tail -1
timeout 120 python tools/profile_target.py 300 layer4.1.conv1 36 1 2 2>&1 | tail -1
timeout 120 python tools/profile_target.py 300 layer2.1.conv1 8 1 0 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py 2>&1 | tail -1
