cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -12
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v16.txt | tail -1
timeout 300 python bench.py --steps 20 2>gpurun_out/bench_v16.err >gpurun_out/bench_v16.json
python -c "import json;d=json.load(open('gpurun_out/bench_v16.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
MCP_BA_FUSE_SCHUR=0 timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1
timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1
timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1
MCP_BA_FUSE_SCHUR=0 timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1
