mkdir -p gpurun_out/r2aa
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2aa/pytest.log 2>&1; tail -3 gpurun_out/r2aa/pytest.log | cut -c1-300
python tools/prof_epc.py 512 512 9 6 2>&1 | tee gpurun_out/r2aa/prof_epc.log | grep -v "^pieces"
python tools/time_epc_info.py 3 2>&1 | tee gpurun_out/r2aa/epc_info.log
python bench.py 2>gpurun_out/r2aa/bench.err | tee gpurun_out/r2aa/bench.json | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac']); print(b['roofline'].get('mttkrp')); print(b['cpu_baseline']['value'], b['reference_eager_b200']['value'], b['parity_mode'])"
