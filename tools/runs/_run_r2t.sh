mkdir -p gpurun_out/r2t
for spec in "layer4.1.conv1 33 0" "layer4.1.conv1 36 0" "layer4.1.conv1 148 0" "layer4.1.conv1 33 2" "layer3.1.conv1 7 0" "layer2.1.conv1 2 0"; do
  set -- $spec
  python tools/profile_target.py 300 $1 $2 1 $3
done
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2t/pytest.log 2>&1; tail -3 gpurun_out/r2t/pytest.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg 2>/dev/null | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac']); print(b['config']['ctas_per_unit'])"
