set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_fe_gpu.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -5
for n in 1 3; do
  MCP_BA_SPECULATE=$n timeout 300 python bench.py --no-cpu-baseline --no-frontend --steps 10 --warmup 3 > gpurun_out/bench_spec$n.json 2> gpurun_out/bench_spec$n.err
  python -c "import json;d=json.load(open('gpurun_out/bench_spec$n.json'));print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d.get('kernel_ms'))"
done
MCP_BA_SELECT_MULTI=1 MCP_BA_SPECULATE=3 timeout 300 python bench.py --no-cpu-baseline --no-frontend --steps 10 --warmup 3 > gpurun_out/bench_selmulti.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_selmulti.json'));print('selmulti', d['value'], d['ms_per_step'])"
