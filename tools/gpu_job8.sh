cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/prof_ba.py cfg2 10 5
MCP_BA_LOOKAHEAD=0 timeout 300 python tools/prof_ba.py cfg2 10 5
