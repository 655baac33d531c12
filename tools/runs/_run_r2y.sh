mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -x -k "als_epc" 2>&1 | tail -4 | cut -c1-400
python tools/time_epc_info.py 0 2 3 2>&1 | tee gpurun_out/r2y/epc_info.log
timeout 600 python bench.py --full --init parafac-epc --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e 2>gpurun_out/r2y/full_r18.err | tee gpurun_out/r2y/full_r18_epc.json | cut -c1-1800
