mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -x -k "tap_factor or early_exit or cluster_loop" > gpurun_out/r2u/tap.log 2>&1; tail -5 gpurun_out/r2u/tap.log | cut -c1-300; grep "\[tap\]" gpurun_out/r2u/tap.log
