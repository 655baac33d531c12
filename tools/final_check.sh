timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.json; cut -c1-200 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-400
