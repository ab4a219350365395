set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/ba_breakdown.py cfg2
