mkdir -p gpurun_out/r2y
python tools/time_epc_info.py 0 2 3 2>&1 | tee gpurun_out/r2y/epc_info.log
python bench.py --workload layer1 --full --init parafac-epc --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e 2>gpurun_out/r2y/full_l1.err | tee gpurun_out/r2y/full_l1_epc.json | cut -c1-1500
timeout 600 python bench.py --full --init parafac-epc --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e 2>gpurun_out/r2y/full_r18.err | tee gpurun_out/r2y/full_r18_epc.json | cut -c1-3000
