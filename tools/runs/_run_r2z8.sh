mkdir -p gpurun_out/r2z8
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg > gpurun_out/r2z8/bench_r18_n8.json 2> gpurun_out/r2z8/bench_r18_n8.err; echo rc=$?; tail -c 1500 gpurun_out/r2z8/bench_r18_n8.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload sweep256 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-reference --no-parity-leg > gpurun_out/r2z8/bench_sweep256_n8.json 2> gpurun_out/r2z8/bench_sweep256_n8.err; echo rc=$?; tail -c 800 gpurun_out/r2z8/bench_sweep256_n8.json
