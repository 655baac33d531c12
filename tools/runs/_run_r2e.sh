mkdir -p gpurun_out/r2e
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "whole_model_driver" > gpurun_out/r2e/driver.log 2>&1; tail -4 gpurun_out/r2e/driver.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; tail -2 gpurun_out/r2e/bench.err
python bench.py --workload sweep256 --steps 1 --warmup 1 > gpurun_out/r2e/sweep256.json 2> gpurun_out/r2e/sweep256.err; tail -2 gpurun_out/r2e/sweep256.err
python bench.py --workload layer1 --full --no-e2e > gpurun_out/r2e/full_layer1.json 2> gpurun_out/r2e/full_layer1.err; tail -2 gpurun_out/r2e/full_layer1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2e/ref.json 2> gpurun_out/r2e/ref.err; tail -2 gpurun_out/r2e/ref.err
