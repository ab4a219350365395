# gpurun command list behind profiles/r02_*: bench (both arms), launch list of the bench command, ncu --set full of every BA
# kernel of a trial round (cfg2), stream timeline.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_capture_r02.sh'
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
DIGEST=$(python -c "import bench; print(bench.csrc_digest())")   # kernel + launcher sources (not the host marshalling)
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frontend > gpurun_out/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_chol_solve|k_pose_blocks|k_linearize|k_backsub_eval|k_schur_vinv_multi|k_schur_pairs_multi_ca|k_select_cluster|k_lm_control|k_robust_sum' -s 40 -c 16 -f -o gpurun_out/r02_prof_ba python tools/prof_ba.py cfg2 10 2 > gpurun_out/r02_prof.log 2>&1; tail -2 gpurun_out/r02_prof.log
(echo "# csrc $DIGEST"; python tools/ncu_extract.py gpurun_out/r02_prof_ba.ncu-rep) > gpurun_out/r02_ncu_full_ba.txt
cp gpurun_out/r02_ncu_full_ba.txt profiles/r02_ncu_full_ba.txt   # bench.py quotes roofline.traffic from the capture of THIS build (digest in line 1)
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 300 ncu --set full --clock-control none -k regex:'k_patch_search|k_fast_score|k_halfsample_fused|k_fast_compact|k_glare_mask' -c 8 -f -o gpurun_out/r02_prof_fe python tools/prof_ba.py tiny 2 1 fe > gpurun_out/r02_prof_fe.log 2>&1
(echo "# csrc $DIGEST"; python tools/ncu_extract.py gpurun_out/r02_prof_fe.ncu-rep) > gpurun_out/r02_ncu_full_fe.txt
MCP_BA_TIMELINE=1 timeout 300 python tools/prof_ba.py cfg2 10 3 2> gpurun_out/r02_timeline_cfg2.txt | tail -1
timeout 200 python tools/ba_breakdown.py cfg2 10 > gpurun_out/r02_breakdown_cfg2.txt 2>&1; tail -4 gpurun_out/r02_breakdown_cfg2.txt
timeout 300 python tools/ba_breakdown.py cfg4 10 > gpurun_out/r02_breakdown_cfg4_1gpu.txt 2>&1; tail -4 gpurun_out/r02_breakdown_cfg4_1gpu.txt
timeout 100 python tools/fe_bench.py 300 | tee gpurun_out/r02_fe_bench.txt | tail -1
timeout 100 python tools/load_bench.py cfg2 | tee gpurun_out/r02_load_bench.txt | tail -3
timeout 100 python tools/e2e_bench.py cfg2 10 40 | tee gpurun_out/r02_e2e_bench.txt | tail -1
ls -la gpurun_out/r02_*
