cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1; timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v20.txt | tail -1
timeout 300 python bench.py --steps 20 2>gpurun_out/bench_v20.err >gpurun_out/bench_v20.json
python -c "import json;d=json.load(open('gpurun_out/bench_v20.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
