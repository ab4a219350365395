cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_schur_rows' -s 10 -c 1 -f -o gpurun_out/prof_rows python tools/prof_ba.py cfg2 10 1 > /dev/null 2>&1
MCP_BA_SCHUR=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_schur_pairs_tma' -s 10 -c 1 -f -o gpurun_out/prof_pairs python tools/prof_ba.py cfg2 10 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
