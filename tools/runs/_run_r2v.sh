mkdir -p gpurun_out/r2v
for B in "" "--budgets 1,1,1,1,1,2,2,2,3,8,8,8,14,32,32,32" "--budgets 1,1,1,1,1,2,2,2,3,8,8,8,13,33,32,32" "--budgets 1,1,1,1,2,2,2,2,4,8,8,8,12,32,32,32"; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e $B 2>/dev/null | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$B', {k:round(b[k],1) for k in ('value','ms_per_step')}, round(b['roofline']['frac'],4)); print(list(b['config']['ctas_per_unit'].values())); print(list(b['per_unit_sweep_ms_last_step'].values()))"
done
