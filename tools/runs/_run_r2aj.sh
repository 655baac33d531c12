mkdir -p gpurun_out/r2aj
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2aj/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aj/pytest.log | cut -c1-200
for g in 33 64 148; do timeout 120 python tools/profile_target.py 300 layer4.1.conv1 $g 1 2; done 2>&1
for g in 33 36 74 148; do timeout 120 python tools/profile_target.py 300 layer4.1.conv1 $g 1 0; done 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg 2>/dev/null | tee gpurun_out/r2aj/bench.json | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac']); print(b['per_unit_sweep_ms_last_step'])"
