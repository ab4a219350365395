set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests/test_ba_multigpu.py -x -q -m gpu 2>&1 | tail -5
MCP_BA_SPECULATE_MULTI=1 timeout 300 $TR tools/ba_multi.py cfg2 10 2>&1 | grep -E "MULTI|Error|error" | head
timeout 300 $TR tools/ba_multi.py cfg2 10 2>&1 | grep -E "MULTI|Error|error" | head
MCP_BA_SPECULATE_MULTI=3 timeout 300 $TR tools/ba_multi.py cfg2 10 2>&1 | grep -E "MULTI|Error|error" | head
MCP_BA_SPECULATE_MULTI=1 timeout 400 $TR tools/ba_multi.py cfg4 10 2>&1 | grep -E "MULTI|Error|error" | head
timeout 300 $TR tools/ba_multi.py cfg4 10 2>&1 | grep -E "MULTI|Error|error" | head
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_v10_2gpu.json 2> gpurun_out/bench_v10_2gpu.err; tail -5 gpurun_out/bench_v10_2gpu.err
python -c "import json;d=json.load(open('gpurun_out/bench_v10_2gpu.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['scale_big_map'])"
timeout 300 python tools/ba_breakdown.py cfg2
timeout 300 python tools/ba_breakdown.py cfg4
MCP_BA_SPECULATE=1 timeout 300 python tools/prof_ba.py cfg4 10 3
MCP_BA_SPECULATE=2 timeout 300 python tools/prof_ba.py cfg4 10 3
