"""Stream timeline (MCP_BA_TIMELINE=1, rank 0) of the point-sharded bundle adjuster under torchrun.
usage: torchrun ... tools/ba_multi_timeline.py [cfg] [lm_iters]   -> TL lines on rank 0's stderr, summary on stdout"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcptam_b200 import capi, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
prob = synth.make_ba_config(cfg, seed=0) if rank == 0 else None
dist.barrier()
if prob is None:
    prob = synth.make_ba_config(cfg, seed=0)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
h = capi.BaHandle(device=lr)
h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
h.load(prob)
for _ in range(2):
    h.reset_state(); rc, st = h.compute(iters)
dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    os.environ["MCP_BA_TIMELINE"] = "1"
h.reset_state(); rc, st = h.compute(iters)
os.environ.pop("MCP_BA_TIMELINE", None)
h.reset_state(); rc, st = h.compute(iters)
if rank == 0:
    print("MULTI_TL", cfg, "world", world, "rc", rc, "trials", st.total_trials, "gpu_ms", st.gpu_ms, "it/s", 1e3 * rc / st.gpu_ms, flush=True)
dist.barrier()
dist.destroy_process_group()
