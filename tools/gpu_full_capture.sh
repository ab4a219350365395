set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --impl reference > gpurun_out/bench_v21_ref.json 2> gpurun_out/bench_v21_ref.err
timeout 600 python bench.py > gpurun_out/bench_v21.json 2> gpurun_out/bench_v21.err; tail -3 gpurun_out/bench_v21.err
python -c "import json;d=json.load(open('gpurun_out/bench_v21.json'));print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['roofline']['per_kernel_ms_per_step'], d['gpu_launches'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v21.csv python tools/prof_ba.py cfg2 10 2 fe > gpurun_out/launches_v21.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_chol_solve|k_pose_blocks|k_linearize|k_backsub_eval|k_schur_vinv_multi|k_schur_pairs_multi_ca|k_select_cluster|k_lm_control' -s 40 -c 12 -f -o gpurun_out/prof_v21_ba python tools/prof_ba.py cfg2 10 2 > gpurun_out/prof_v21.log 2>&1; tail -2 gpurun_out/prof_v21.log
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/timeline_v21.txt | tail -1
ls -la gpurun_out/*v21*
