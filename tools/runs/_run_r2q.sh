mkdir -p gpurun_out/r2q
for cfg in "16 0" "16 1" "32 0" "8 0"; do
  set -- $cfg
  if [ $2 = 1 ]; then export ADMMQ_NO_CLUSTER=1; else unset ADMMQ_NO_CLUSTER; fi
  python bench.py --workload sweep256 --steps 1 --warmup 1 --no-e2e --round-size $1 2>/dev/null | python -c "
import json,sys; b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('round-size $1 no_cluster $2:', round(b['value']), round(b['ms_per_step'],1))"
done
