set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -3 gpurun_out/bench_v9.err
python -c "import json;d=json.load(open('gpurun_out/bench_v9.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['per_kernel_ms_per_step'], d['roofline']['per_kernel_launches_per_step'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_v9_ref.json 2>&1; cat gpurun_out/bench_v9_ref.json | cut -c1-300
MCP_BA_SPECULATE=0 timeout 300 python bench.py --no-cpu-baseline --no-frontend --steps 20 > gpurun_out/bench_v9_nospec.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_v9_nospec.json'));print('nospec', d['value'], d['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v9.csv python tools/prof_ba.py cfg2 10 2 fe > gpurun_out/launches_v9.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_v9.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_linearize|k_chol_solve|k_schur_pairs|k_backsub|k_schur_y' -s 30 -c 8 -f -o gpurun_out/prof_v9 python tools/prof_ba.py cfg2 10 1 > gpurun_out/prof_v9.log 2>&1; tail -3 gpurun_out/prof_v9.log
timeout 300 python tools/prof_ba.py cfg4 10 3
ls -la gpurun_out
