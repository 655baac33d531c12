mkdir -p gpurun_out/r2m
python tools/time_mttkrp.py 2>&1 | cut -c1-330
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_mttkrp or contractions or outer_loop_with_tensor" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err; tail -3 gpurun_out/r2m/bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2m/bench.json') if l.startswith('{')][-1])
print({k:b[k] for k in ('value','ms_per_step','gpu_launches')}, b['e2e']['value'], b['roofline']['frac'])
print('parity', b['parity_mode']); print('cpu', b['cpu_baseline']['value'], b['cpu_baseline']['cores'], 'eager', b['reference_eager_b200'].get('value'))
PY
