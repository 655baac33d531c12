timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1g_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_admm_loop -s 1 -c 1 -o gpurun_out/r1g_admm_loop_l4_36ctas python tools/profile_target.py 100 layer4.1.conv1 36 1 0 > gpurun_out/r1g_ncu_m0.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_admm_loop -s 1 -c 1 -o gpurun_out/r1g_admm_loop_l4_skinny python tools/profile_target.py 100 layer4.1.conv1 36 1 2 > gpurun_out/r1g_ncu_m2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mse_sums -s 2 -c 1 -o gpurun_out/r1g_search_512x1141 python tools/profile_search.py 512 1141 4 200 33 1 > gpurun_out/r1g_ncu_s.log 2>&1
tail -2 gpurun_out/r1g_ncu_m0.log
