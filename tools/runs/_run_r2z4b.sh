mkdir -p gpurun_out/r2z4b
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2z4b/bench_r18_n4.json 2> gpurun_out/r2z4b/bench_r18_n4.err; echo rc=$?
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2z4b/bench_r18_n2.json 2> gpurun_out/r2z4b/bench_r18_n2.err; echo rc=$?
python - <<'PY'
import json
for n in (4, 2):
    b = json.loads([l for l in open(f'gpurun_out/r2z4b/bench_r18_n{n}.json') if l.startswith('{')][-1])
    print(n, {k: b.get(k) for k in ('value', 'ms_per_step', 'factor_gather_ms', 'per_rank_ms_per_step')}, 'e2e', b['e2e']['ms_per_step'], b['config']['units_per_rank'])
PY
