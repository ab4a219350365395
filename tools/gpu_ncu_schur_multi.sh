cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_schur_pairs_multi|k_schur_vinv' -s 10 -c 2 -f -o gpurun_out/prof_multi_ca python tools/prof_ba.py cfg2 10 1 > /dev/null 2>&1
MCP_BA_SCHUR_STAGE=tma timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_schur_pairs_multi' -s 5 -c 1 -f -o gpurun_out/prof_multi_tma python tools/prof_ba.py cfg2 10 1 > /dev/null 2>&1
ls -la gpurun_out/prof_multi*
