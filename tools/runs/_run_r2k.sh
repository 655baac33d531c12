mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2k/pytest.log 2>&1; tail -6 gpurun_out/r2k/pytest.log | cut -c1-300
grep "cluster\]\|outer-full\]\|outer-full-2\|als-epc\|driver\]" gpurun_out/r2k/pytest.log | cut -c1-330
python tools/time_mttkrp.py 2>&1 | cut -c1-250
python bench.py --workload layer1 --steps 3 --warmup 2 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('layer1', b['value'], b['ms_per_step'], b['e2e']['value'])"
