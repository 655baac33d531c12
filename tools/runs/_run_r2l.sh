mkdir -p gpurun_out/r2l
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2l/pytest.log 2>&1; tail -8 gpurun_out/r2l/pytest.log | cut -c1-300
grep "first-N/999\|teacher-forced late\|top-1\]" gpurun_out/r2l/pytest.log | cut -c1-400
