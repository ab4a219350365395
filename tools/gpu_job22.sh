cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
for m in tma8 tma16 ca; do echo $m; MCP_BA_SCHUR_STAGE=$m timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1; MCP_BA_SCHUR_STAGE=$m timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1; done
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v18.txt | tail -1
