cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
MCP_BA_LOAD_TRACE=1 timeout 600 python bench.py --steps 20 > gpurun_out/bench_v13b.json 2> gpurun_out/bench_v13b.err; grep mcp_ba_load gpurun_out/bench_v13b.err | tail -4
python -c "import json;d=json.load(open('gpurun_out/bench_v13b.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
lscpu | grep -E "Model name|^CPU\(s\)|MHz"
