"""Multi-GPU check of the point-sharded bundle adjuster (run under torchrun, one rank per GPU):
sharded result == single-GPU result (sum order differs: 1e-9 relative), and timing of both."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcptam_b200 import capi, synth  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    prob = synth.make_ba_config(cfg, seed=0) if rank == 0 else None
    dist.barrier()
    if prob is None:
        prob = synth.make_ba_config(cfg, seed=0)          # disk cache written by rank 0
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    h = capi.BaHandle(device=lr)
    h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    h.load(prob)
    for _ in range(2):
        h.reset_state()
        rc, st = h.compute(iters)
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    h.reset_state()
    rc, st = h.compute(iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    P, X = h.poses(), h.points()
    if rank == 0:
        s = capi.BaHandle(device=lr)
        s.load(prob)
        for _ in range(2):
            s.reset_state()
            rc1, st1 = s.compute(iters)
        t = time.perf_counter()
        s.reset_state()
        rc1, st1 = s.compute(iters)
        dt1 = time.perf_counter() - t
        relp = np.linalg.norm(P - s.poses()) / np.linalg.norm(s.poses())
        relx = np.linalg.norm(X - s.points()) / np.linalg.norm(s.points())
        ok = rc == rc1 and st.total_trials == st1.total_trials and relp < 1e-7 and relx < 1e-7
        print("MULTI", cfg, "world", world, "rc", rc, rc1, "trials", st.total_trials, st1.total_trials, "rel", relp, relx,
              "gpu_ms sharded", st.gpu_ms, "single", st1.gpu_ms, "wall", dt, dt1, "OK" if ok else "MISMATCH", flush=True)
    # the abort flag of ONE rank must stop every rank in the same round (the decision is collective: no hang, same return code)
    ab = np.array([1 if rank == world - 1 else 0], np.uint8)
    h.reset_state()
    rc_a, st_a = h.compute(iters, abort=ab)
    rcs = [None] * world
    dist.all_gather_object(rcs, (rc_a, st_a.iterations))
    if rank == 0:
        print("MULTI_ABORT", rcs, "OK" if all(r == rcs[0] for r in rcs) and rcs[0][0] == 0 else "MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
