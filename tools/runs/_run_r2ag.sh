mkdir -p gpurun_out/r2ag
timeout 300 python bench.py --workload layer1 --full --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2ag/full_l1_random.json
timeout 300 python bench.py --workload layer1 --no-cpu-baseline --no-eager-reference --no-parity-leg 2>/dev/null | tail -1 > gpurun_out/r2ag/bench_l1.json
timeout 400 python bench.py --full --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2ag/full_r18_random.json
python - <<'PY'
import json
for f in ('full_l1_random', 'full_r18_random'):
    b = json.load(open(f'gpurun_out/r2ag/{f}.json'))
    print(f, b['value'], 'init', b.get('per_rank_init_wall_s'), 'iters', b.get('inner_iterations'), 'sweeps', sorted(b['sweeps'].values())[:1], sorted(b['sweeps'].values())[-1:])
b = json.load(open('gpurun_out/r2ag/bench_l1.json'))
print('layer1', b['value'], b['ms_per_step'], b['e2e']['value'])
PY
