cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_multigpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_v14_2gpu.json 2> gpurun_out/bench_v14_2gpu.err; tail -3 gpurun_out/bench_v14_2gpu.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_v14_2gpu.json') if l.startswith('{')][-1]);print(d['value'], d['ms_per_step'], d['e2e'], d['scale_big_map'])"
