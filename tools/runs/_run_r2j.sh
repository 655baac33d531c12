mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2j/pytest.log 2>&1; tail -6 gpurun_out/r2j/pytest.log | cut -c1-300
grep "cluster\]\|first-N\|teacher-forced late" gpurun_out/r2j/pytest.log | cut -c1-300
python tools/time_mttkrp.py 2>&1 | head -3 | cut -c1-250
python bench.py --workload layer1 --steps 3 --warmup 2 --no-cpu-baseline --no-eager-reference > gpurun_out/r2j/bench_layer1.json 2> gpurun_out/r2j/bench_layer1.err; tail -2 gpurun_out/r2j/bench_layer1.err; python -c "
import json; b=json.load(open('gpurun_out/r2j/bench_layer1.json')); print('layer1', b['value'], b['ms_per_step'], b['e2e']['value'], b['config']['ctas_per_unit'])"
python bench.py --workload layer1 --full --no-e2e 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full layer1 random', b['value'], b['sweeps'], b['rec_error'])"
