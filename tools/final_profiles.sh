# end-of-round ncu captures (run on the GPU box after the un-profiled commands have exited 0)
mkdir -p gpurun_out/r2p
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2p/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-eager-reference > gpurun_out/r2p/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_admm_loop -s 1 -c 1 -o gpurun_out/r2p/r2_admm_loop_l4_36ctas python tools/profile_target.py 100 layer4.1.conv1 36 1 0 > gpurun_out/r2p/ncu_m0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_admm_loop_cluster -s 1 -c 1 -o gpurun_out/r2p/r2_admm_loop_cluster_l1 python tools/profile_target.py 100 layer1.0.conv1 8 1 0 > gpurun_out/r2p/ncu_cl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mttkrp_fold_tc -s 2 -c 1 -o gpurun_out/r2p/r2_mttkrp_fold python tools/time_mttkrp.py > gpurun_out/r2p/ncu_mt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mttkrp_foldlong_tc -s 2 -c 1 -o gpurun_out/r2p/r2_mttkrp_foldlong python tools/time_mttkrp.py > gpurun_out/r2p/ncu_mt2.log 2>&1
tail -2 gpurun_out/r2p/ncu_m0.log gpurun_out/r2p/ncu_cl.log gpurun_out/r2p/ncu_mt.log
ls -la gpurun_out/r2p
