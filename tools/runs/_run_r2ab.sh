mkdir -p gpurun_out/r2ab
nproc; uptime
timeout 400 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "model_top1" --durations=5 -o faulthandler_timeout=200 > gpurun_out/r2ab/top1.log 2>&1; echo rc=$?; tail -25 gpurun_out/r2ab/top1.log | cut -c1-400
uptime
