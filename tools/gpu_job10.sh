cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 120 python tools/ba_breakdown.py cfg2 2>&1 | grep -E "it/s|profiled"
MCP_BA_SCHUR=1 timeout 120 python tools/ba_breakdown.py cfg2 2>&1 | grep -E "it/s|profiled"
