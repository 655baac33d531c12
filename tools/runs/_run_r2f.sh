mkdir -p gpurun_out/r2f
nvidia-smi -L > gpurun_out/r2f/gpus.txt
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "whole_model_driver" > gpurun_out/r2f/driver.log 2>&1; tail -4 gpurun_out/r2f/driver.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f/bench_n2.json 2> gpurun_out/r2f/bench_n2.err; tail -2 gpurun_out/r2f/bench_n2.err
$TR bench.py --gpus 2 --workload sweep256 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2f/sweep256_n2.json 2> gpurun_out/r2f/sweep256_n2.err; tail -2 gpurun_out/r2f/sweep256_n2.err
python bench.py --steps 2 --warmup 3 --cpu-budget-s 6 > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err; tail -2 gpurun_out/r2f/bench_n1.err
