cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; tail -2 gpurun_out/bench_v22.err
python -c "import json;d=json.load(open('gpurun_out/bench_v22.json'));print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])"
