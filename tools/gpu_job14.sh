set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_host_gpu.py tests/test_stream_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; tail -3 gpurun_out/bench_v13.err
python -c "import json;d=json.load(open('gpurun_out/bench_v13.json'));print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
