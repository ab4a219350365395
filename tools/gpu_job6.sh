cd $GRAFT_REPO_ROOT
for v in 0 1 2 4 7; do echo "experiment $v"; MCP_BA_EXPERIMENT=$v timeout 300 python tools/ba_breakdown.py cfg2 2 2>&1 | grep -E "profiled"; done
