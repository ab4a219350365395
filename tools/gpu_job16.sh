cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -3
for th in 1 4 8; do
MCP_BA_HOST_THREADS=$th MCP_BA_LOAD_TRACE=1 timeout 600 python bench.py --steps 20 > gpurun_out/bench_v14_t$th.json 2> gpurun_out/bench_v14_t$th.err; grep mcp_ba_load gpurun_out/bench_v14_t$th.err | tail -2
python -c "import json;d=json.load(open('gpurun_out/bench_v14_t$th.json'));print($th, d['value'], d['ms_per_step'], d['e2e'])"
done
