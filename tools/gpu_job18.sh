cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -2
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v15.txt | tail -1
MCP_BA_LOAD_TRACE=1 timeout 300 python bench.py --steps 20 2>&1 >gpurun_out/bench_v15.json | grep mcp_ba_load | tail -1
python -c "import json;d=json.load(open('gpurun_out/bench_v15.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
