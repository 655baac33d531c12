mkdir -p gpurun_out/r2z4
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg > gpurun_out/r2z4/bench_r18_n4.json 2> gpurun_out/r2z4/bench_r18_n4.err; echo rc=$?; tail -c 1800 gpurun_out/r2z4/bench_r18_n4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --workload sweep256 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-reference --no-parity-leg > gpurun_out/r2z4/bench_sweep256_n4.json 2> gpurun_out/r2z4/bench_sweep256_n4.err; echo rc=$?; tail -c 1200 gpurun_out/r2z4/bench_sweep256_n4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
