mkdir -p gpurun_out/r2ae
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ae/pytest.log | cut -c1-300
for spec in "layer4.1.conv1 33 0" "layer4.1.conv1 36 0" "layer4.1.conv1 33 2" "layer3.1.conv1 7 0" "layer2.1.conv1 2 0"; do
  set -- $spec
  timeout 120 python tools/profile_target.py 300 $1 $2 1 $3
done 2>&1 | tee gpurun_out/r2ae/phases.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg 2>/dev/null | tee gpurun_out/r2ae/bench.json | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac']); print(b['per_unit_sweep_ms_last_step'])"
