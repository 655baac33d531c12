timeout 120 python tools/profile_target.py 300 layer3.1.conv1 7 1 0 2>&1 | tail -1
timeout 120 python tools/profile_target.py 300 layer4.1.conv1 36 1 0 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>&1 | tail -1
