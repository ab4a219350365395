set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
tools/_build/potrf_bench
timeout 600 python -m pytest tests/test_fe_gpu.py -x -q -m gpu -k "rest" 2>&1 | tail -15
