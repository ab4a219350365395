set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-frontend --steps 10 --warmup 3 > gpurun_out/bench_sel5.json 2> gpurun_out/bench_sel5.err
python -c "import json;d=json.load(open('gpurun_out/bench_sel5.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_kernel_ms_per_step'])"
