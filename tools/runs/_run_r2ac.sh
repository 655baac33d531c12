mkdir -p gpurun_out/r2ac
for i in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-reference --no-parity-leg --no-e2e --placement 2>gpurun_out/r2ac/place$i.log | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print({k:round(b[k],1) for k in ('value','ms_per_step')})"
grep placement gpurun_out/r2ac/place$i.log | cut -c1-700
done
