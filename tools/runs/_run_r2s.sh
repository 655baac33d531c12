mkdir -p gpurun_out/r2s
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2s/pytest.log 2>&1; tail -5 gpurun_out/r2s/pytest.log | cut -c1-300
python tools/time_mttkrp.py 2>&1 | head -3 | cut -c1-330
python bench.py --steps 3 --warmup 3 --cpu-budget-s 8 > gpurun_out/r2s/bench.json 2> gpurun_out/r2s/bench.err; tail -3 gpurun_out/r2s/bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2s/bench.json') if l.startswith('{')][-1])
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac'], b['parity_mode'])
print(b['config']['ctas_per_unit']); print(b['per_unit_sweep_ms_last_step'])
PY
