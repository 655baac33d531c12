mkdir -p gpurun_out/r2an
for w in resnet50-l4 llama7b; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 2 --no-eager-reference --no-parity-leg --cpu-budget-s 6 2>gpurun_out/r2an/bench_$w.err | tail -1 > gpurun_out/r2an/bench_$w.json
  python - <<PY
import json
b = json.load(open('gpurun_out/r2an/bench_$w.json'))
print('$w', {k: b[k] for k in ('value', 'ms_per_step')}, 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'], 'cpu', b['cpu_baseline']['value'])
PY
done
