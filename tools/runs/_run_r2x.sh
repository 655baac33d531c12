mkdir -p gpurun_out/r2x
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -x -k "als_epc" > gpurun_out/r2x/epc_test.log 2>&1; tail -4 gpurun_out/r2x/epc_test.log | cut -c1-400
(python tools/prof_epc.py 64 64 9 20; python tools/prof_epc.py 512 512 9 8) 2>&1 | tee gpurun_out/r2x/prof_epc.log
timeout 700 python tools/time_init.py 4 2>&1 | tee gpurun_out/r2x/time_init.log
