# end-of-round checks on the GPU box: GPU suite, smoke(), default bench line, reference arm (every step under its own timeout)
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/final/pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/final/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/final/bench.err | tail -1 > gpurun_out/final/bench.json; echo "bench rc=$?"
python - <<'PY'
import json
b = json.load(open('gpurun_out/final/bench.json'))
print({k: b[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'])
print('mttkrp', b['roofline'].get('mttkrp'))
print('cpu', b['cpu_baseline']['value'], 'eager', b['reference_eager_b200']['value'], 'parity', b['parity_mode']['ms_per_step'], b['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final/bench_reference.json; cut -c1-300 gpurun_out/final/bench_reference.json
