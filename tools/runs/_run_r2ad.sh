mkdir -p gpurun_out/r2ad
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "als_epc" 2>&1 | tail -2
for v in 0 1 2; do
  echo "ADMMQ_MTTKRP_F64=$v"
  ADMMQ_MTTKRP_F64=$v timeout 200 python tools/prof_epc.py 512 512 9 4 2>&1 | grep "^mode" | cut -c1-140
  ADMMQ_MTTKRP_F64=$v timeout 200 python tools/prof_epc.py 256 256 9 4 2>&1 | grep "^mode" | cut -c1-140
done 2>&1 | tee gpurun_out/r2ad/mttkrp_f64_variants.log
for w in resnet50-l4 llama7b; do
  timeout 400 python bench.py --workload $w --steps 2 --warmup 2 --no-eager-reference --no-parity-leg --cpu-budget-s 8 2>gpurun_out/r2ad/bench_$w.err | tail -1 > gpurun_out/r2ad/bench_$w.json
  python - <<PY
import json
b = json.load(open('gpurun_out/r2ad/bench_$w.json'))
print('$w', {k: b[k] for k in ('value', 'ms_per_step')}, 'e2e', b['e2e']['value'] if b.get('e2e') else None, 'frac', (b.get('roofline') or {}).get('frac'), 'cpu', (b.get('cpu_baseline') or {}).get('value'))
print(b['config']['workload'][:200])
PY
done
