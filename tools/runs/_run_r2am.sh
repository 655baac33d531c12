timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python tools/time_epc_info.py 0 2>&1 | tail -1 | cut -c1-250
