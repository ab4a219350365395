cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -12
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v17.txt | tail -1
echo ca; timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1
echo tma; MCP_BA_SCHUR_STAGE=tma timeout 300 python tools/prof_ba.py cfg2 10 5 | tail -1
echo ca4; timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1
echo tma4; MCP_BA_SCHUR_STAGE=tma timeout 300 python tools/prof_ba.py cfg4 10 3 | tail -1
timeout 300 python bench.py --steps 20 2>gpurun_out/bench_v17.err >gpurun_out/bench_v17.json
python -c "import json;d=json.load(open('gpurun_out/bench_v17.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
