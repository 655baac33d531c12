mkdir -p gpurun_out/r2o
nvidia-smi -L | wc -l
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N"
  $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2o/bench_n$N.json 2> gpurun_out/r2o/bench_n$N.err; grep -v "OMP_NUM\|^\*" gpurun_out/r2o/bench_n$N.err | tail -3
  $TR bench.py --gpus $N --workload sweep256 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2o/sweep256_n$N.json 2> gpurun_out/r2o/sweep256_n$N.err; grep -v "OMP_NUM\|^\*" gpurun_out/r2o/sweep256_n$N.err | tail -3
done
python - <<'PY'
import json
for f in ('bench_n8','bench_n4','sweep256_n8','sweep256_n4'):
    try:
        b=json.loads([l for l in open(f'gpurun_out/r2o/{f}.json') if l.startswith('{')][-1])
        print(f, round(b['value']), round(b['ms_per_step'],1), b['per_rank_ms_per_step'], 'gather', b['factor_gather_ms'], b['config'].get('units_per_rank'))
    except Exception as e: print(f, 'failed', e)
PY
