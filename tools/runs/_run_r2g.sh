mkdir -p gpurun_out/r2g
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "als_epc" > gpurun_out/r2g/alsepc.log 2>&1; tail -6 gpurun_out/r2g/alsepc.log | cut -c1-700
python tools/prof_epc.py 64 64 9 20 > gpurun_out/r2g/prof_l1.log 2>&1; cat gpurun_out/r2g/prof_l1.log
python tools/prof_epc.py 512 512 9 5 > gpurun_out/r2g/prof_l4.log 2>&1; cat gpurun_out/r2g/prof_l4.log
python tools/prof_epc.py 256 256 9 5 > gpurun_out/r2g/prof_l3.log 2>&1; cat gpurun_out/r2g/prof_l3.log
python bench.py --workload layer1 --full --init parafac-epc --no-e2e > gpurun_out/r2g/full_layer1_epc.json 2> gpurun_out/r2g/full_layer1_epc.err; tail -c 1200 gpurun_out/r2g/full_layer1_epc.json; tail -3 gpurun_out/r2g/full_layer1_epc.err
