mkdir -p gpurun_out/r2r
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2r/pytest.log 2>&1; tail -5 gpurun_out/r2r/pytest.log | cut -c1-300
for spec in "layer4.1.conv1 32 0" "layer4.1.conv1 33 0" "layer4.1.conv1 34 0" "layer4.1.conv1 36 0" "layer4.1.conv1 148 0" "layer3.1.conv1 7 0" "layer3.1.conv1 6 0" "layer4.0.conv1 13 0"; do
  set -- $spec
  python tools/profile_target.py 300 $1 $2 1 $3
done
echo "forced bn=80 at 34:"; ADMMQ_LOOP_BN=80 python tools/profile_target.py 300 layer4.1.conv1 34 1 0
python tools/time_mttkrp.py 2>&1 | head -3 | cut -c1-330
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg > gpurun_out/r2r/bench.json 2> gpurun_out/r2r/bench.err; tail -3 gpurun_out/r2r/bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2r/bench.json') if l.startswith('{')][-1])
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac'])
print(b['config']['ctas_per_unit']); print(b['per_unit_sweep_ms_last_step'])
PY
