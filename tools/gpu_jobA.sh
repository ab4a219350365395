set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; tail -3 gpurun_out/bench_v11.err
python -c "import json;d=json.load(open('gpurun_out/bench_v11.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['per_kernel_ms_per_step'], d['cpu_baseline'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_linearize' -s 3 -c 1 -f -o gpurun_out/prof_v11_lin python tools/prof_ba.py cfg2 10 1 > gpurun_out/prof_v11.log 2>&1; tail -2 gpurun_out/prof_v11.log
