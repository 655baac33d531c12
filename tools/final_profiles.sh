# end-of-round ncu captures (run on the GPU box after the un-profiled commands have exited 0); outputs under gpurun_out/r2q
mkdir -p gpurun_out/r2q
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-eager-reference --no-parity-leg > /dev/null 2>&1; echo "plain bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2q/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-eager-reference --no-parity-leg > gpurun_out/r2q/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 120 python tools/profile_target.py 100 layer4.1.conv1 36 1 0; echo "plain target rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_admm_loop -s 1 -c 1 -o gpurun_out/r2q/r2f_admm_loop_l4_36ctas python tools/profile_target.py 100 layer4.1.conv1 36 1 0 > gpurun_out/r2q/ncu_m0.log 2>&1; echo "full capture rc=$?"
tail -2 gpurun_out/r2q/ncu_m0.log
ls -la gpurun_out/r2q
