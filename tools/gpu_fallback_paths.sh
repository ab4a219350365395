# the selectable non-default paths must stay parity-green: usage: gpurun --timeout 900 -- 'bash tools/gpu_fallback_paths.sh'
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== FE: four-kernel FAST, staged copies, no TMA"; MCP_FE_FAST_FUSED=0 MCP_FE_ZEROCOPY=0 MCP_FE_TMA=0 timeout 300 python -m pytest tests/test_fe_gpu.py -m gpu -q -x 2>&1 | tail -1
echo "== FE: default"; timeout 300 python -m pytest tests/test_fe_gpu.py -m gpu -q -x 2>&1 | tail -1
echo "== BA: no speculative sigma, no folded records"; MCP_BA_SPEC_SIGMA=0 MCP_BA_FOLD_VINV=0 timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "not cfg4" 2>&1 | tail -1
echo "== BA: no look-ahead, no PDL, linearize variant 0"; MCP_BA_LOOKAHEAD=0 MCP_BA_PDL=0 MCP_BA_LIN_VARIANT=0 timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "not cfg4" 2>&1 | tail -1
echo "== BA: one candidate, linearize variant 2"; MCP_BA_SPECULATE=1 MCP_BA_LIN_VARIANT=2 timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k "not cfg4" 2>&1 | tail -1
echo "== load trace (warm, 16 threads: the last loads of the sweep)"; MCP_BA_LOAD_TRACE=1 MCP_PREP_TRACE=1 timeout 100 python tools/load_bench.py cfg2 2>&1 | grep "PREP\|mcp_ba_load" | tail -13
