set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -5
for v in 0 1 2; do MCP_BA_LIN_VARIANT=$v timeout 300 python tools/ba_breakdown.py cfg2 2>&1 | grep -E "it/s|profiled"; done
