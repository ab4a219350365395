cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
MCP_BA_PDL=1 timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
echo pdl0; MCP_BA_PDL=0 timeout 200 python tools/prof_ba.py cfg2 10 5 | tail -1
echo pdl1; MCP_BA_PDL=1 timeout 200 python tools/prof_ba.py cfg2 10 5 | tail -1
echo pdl1-cfg4; MCP_BA_PDL=1 timeout 200 python tools/prof_ba.py cfg4 10 3 | tail -1
MCP_BA_PDL=1 MCP_BA_TIMELINE=1 timeout 200 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v22_pdl.txt | tail -1
