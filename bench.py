#!/usr/bin/env python
"""bench.py — ChainBundle LM iterations/sec on the 4-cam 200 KF / 10k-point synthetic map (BASELINE.json configs[1]).

One "step" = one BundleAdjust-equivalent call: restore the initial estimate, run `--lm-iters` outer LM
iterations (default 10 = the reference's first AdjustAndUpdate(...,10) pass, src/BundleAdjusterMulti.cc:212)
of the ChainBundle bundle adjuster through the C ABI.  `value` = LM iterations / device time with the map
resident in HBM; `e2e` = the same through mcp_ba_load (host buffers -> device) + compute + read-back.
`--impl reference` times the CPU restatement of the reference (oracle/, the reference binary cannot be built
here: ROS/TooN/libCVD/g2o/SuiteSparse are absent) on the host cores.

Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the oracle is only TIMED here (cpu_baseline, --impl reference), never used as a checker: take its -O3 -march=native
# build, compiled on this host (oracle/oracle.py)
os.environ.setdefault("MCP_ORACLE_FAST", "1")

METRIC = "chainbundle_lm_iters_per_sec"
UNIT = "LM iterations/s"
WORKLOAD = "cfg2: 4-cam 200 KF (50 MKF) / 10k-point ChainBundle BA, fp64, seed 0"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, csv line)
        self.windows = []        # [t0, t1] perf_counter intervals of the timed regions

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        # keep the samples taken inside the timed regions (a sample describes the 25 ms before it arrived)
        inside = [ln for t, ln in self.lines if any(a <= t <= b + 0.05 for a, b in self.windows)]
        scope = "timed regions"
        if len(inside) < 3:
            inside, scope = [ln for _, ln in self.lines], "whole run (timed regions shorter than the sampling period)"
        self.scope = scope
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "scope": scope}


def cpu_reference_run(prob, steps, warmup, lm_iters_cpu):
    """Times the CPU restatement (oracle) on the host: one thread, as the reference runs BA on the single
    MapMaker thread (no `#pragma omp` in the reference tree)."""
    from oracle.oracle import OracleBA
    o = OracleBA(prob)
    p0, x0 = np.array(prob.pose_Rt), np.array(prob.pt_xyz)
    total_it, total_t, per_step = 0, 0.0, []
    for s in range(warmup + steps):
        o.set_state(p0, x0)
        t = time.perf_counter()
        rc, st = o.compute(lm_iters_cpu)
        dt = time.perf_counter() - t
        if s >= warmup:
            total_it += max(rc, 0); total_t += dt; per_step.append(dt)
    return total_it / total_t, float(np.mean(per_step)) * 1e3, total_it


def bench_frontend(capi, synth, device, steps=20, warmup=3, n_cams=4, n_patches=1000):
    """BASELINE.json configs[2]: 4-cam 640x480 pyramids, FAST-10 + 1k PatchFinder searches per frame per camera.
    One handle (= one CUDA stream) per camera; host buffers in, host results out (H2D/D2H inside the timing)."""
    import time as _t
    rng = np.random.default_rng(0)
    cams, frames, reqs = [], [], []
    for c in range(n_cams):
        f = capi.FeHandle(640, 480, device=device, max_corners_per_level=16384)
        a = synth.make_frame(seed=100 + c)
        b = synth.make_frame(seed=100 + c, shift=(3.0, -2.0))
        lva = f.make_keyframe(0, a)                         # source keyframe (map)
        cor = lva[0]["corners"]
        cor = cor[(cor[:, 0] > 16) & (cor[:, 0] < 624) & (cor[:, 1] > 16) & (cor[:, 1] < 464)]
        cor = cor[rng.choice(len(cor), n_patches, replace=len(cor) < n_patches)]
        rq = np.zeros(n_patches, capi.PATCH_REQ_DTYPE)
        rq["src_kf"] = 0; rq["src_level"] = 0; rq["src_cx"] = cor[:, 0]; rq["src_cy"] = cor[:, 1]
        rq["warp_inv"] = np.array([1.0, 0.02, -0.02, 1.0]); rq["search_level"] = 0
        rq["pred_x"] = cor[:, 0] - 3 + rng.integers(-2, 3, n_patches); rq["pred_y"] = cor[:, 1] + 2 + rng.integers(-2, 3, n_patches)
        rq["range"] = 10; rq["subpix_its"] = 8
        cams.append(f); frames.append(b); reqs.append(rq)
    t_kf = t_ps = 0.0
    dev_kf = dev_ps = 0.0
    found = 0
    for s in range(warmup + steps):
        for c in range(n_cams):
            t0 = _t.perf_counter()
            cams[c].make_keyframe(1, frames[c])
            t1 = _t.perf_counter()
            res = cams[c].search_patches(1, reqs[c])
            t2 = _t.perf_counter()
            if s >= warmup:
                t_kf += t1 - t0; t_ps += t2 - t1
                tm = cams[c].timing()
                dev_ps += tm["ms_search"]
                found += int(res["found"].sum())
        if s >= warmup:
            pass
    n_frames = steps * n_cams
    tm = cams[0].timing()
    # CPU restatement of the same per-frame work on one host core (the reference tracker is single threaded)
    from oracle import oracle as ora
    t0 = _t.perf_counter()
    pyr_b = ora.pyramid(frames[0])
    lv_b = [ora.level_corners(im) for im in pyr_b]
    t1 = _t.perf_counter()
    pyr_a = ora.pyramid(synth.make_frame(seed=100))
    n_cpu = len(reqs[0])
    ora.search_patches_batch(pyr_a, pyr_b, lv_b, reqs[0][:8])      # warm the marshalling path
    t1b = _t.perf_counter()
    flags_cpu, _ = ora.search_patches_batch(pyr_a, pyr_b, lv_b, reqs[0])
    t2 = _t.perf_counter()
    t1 = t1 + 0.0; t2 = t1 + (t2 - t1b)
    cpu = {"keyframe_ms": 1e3 * (t1 - t0), "patches_per_sec": n_cpu / (t2 - t1), "cores": 1, "kind": "port",
           "sample": "1 camera frame (pyramid + FAST-10 + threshold + LUT) and %d patch searches in one C call, oracle/fe_oracle.c" % n_cpu}
    return {"cpu_baseline": cpu,"workload": "cfg3: %d-cam 640x480, 4-level pyramid + FAST-10 + %d PatchFinder searches (8 sub-pixel its) per frame" % (n_cams, n_patches),
            "camera_frames_per_sec_e2e": n_frames / (t_kf + t_ps), "keyframe_ms_e2e": 1e3 * t_kf / n_frames,
            "patches_per_sec_e2e": n_frames * n_patches / t_ps, "patches_per_sec_device": n_frames * n_patches / (dev_ps * 1e-3),
            "patch_search_ms_device": dev_ps / n_frames, "found_fraction": found / (n_frames * n_patches),
            "pyramid_ms_device": tm["ms_pyramid"], "fast_ms_device": tm["ms_fast"],
            "algorithmic_bytes_per_camera_frame": 307200 + 100800 + 100800, "patch_bytes_each": 800}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lm-iters", type=int, default=10)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frontend", action="store_true")
    ap.add_argument("--scale-config", default="cfg4", help="map used for the extra multi-GPU strong-scaling section (none to skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    from mcptam_b200 import synth
    prob = synth.make_ba_config(args.config, seed=args.seed)
    workload = WORKLOAD if args.config == "cfg2" else args.config

    if args.impl == "reference":
        if rank != 0:
            return 0
        ncores = os.cpu_count()
        lm_cpu = min(args.lm_iters, 3)          # bounded sample: ~1-2 s of CPU work per step
        val, ms, n_it = cpu_reference_run(prob, args.steps, min(warmup, 1), lm_cpu)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(warmup, 1),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "n_pose": prob.n_pose, "n_points": prob.n_pt, "n_meas": prob.n_meas,
                           "lm_iters_per_step": lm_cpu},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores": ncores, "kind": "port",
                                 "sample": "%d LM iterations per step of the same map, CPU restatement (oracle/ba_oracle.c, "
                                           "-O3 -march=native, Schur solve), 1 thread like the reference's MapMaker thread; reference binary "
                                           "unavailable (no ROS/TooN/g2o/SuiteSparse)" % lm_cpu},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / debug lines must not share stdout with the JSON line
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mcptam_b200 import capi

    h = capi.BaHandle(device=local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    h.load(prob)
    stream = torch.cuda.ExternalStream(h.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-data arm: K timed steps, CUDA events on the handle's stream ------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                            # nvidia-smi needs ~1 s to start: launched before the warm-up
    times, iters, launches = [], 0, 0
    for s in range(warmup + args.steps):
        if s == warmup:
            barrier()
            t_wall0 = time.perf_counter()
        flush.fill_(s & 0xFF)                      # L2 flush between steps (outside the timed events)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.reset_state()
        e0.record(stream)
        rc, st = h.compute(args.lm_iters)
        e1.record(stream)
        e1.synchronize()
        if rc <= 0:
            raise SystemExit("bench.py: BA failed rc=%d" % rc)
        if s >= warmup:
            times.append(e0.elapsed_time(e1)); iters += rc; launches += st.kernel_launches
    barrier()
    wall = time.perf_counter() - t_wall0
    sampler.window(t_wall0, t_wall0 + wall)
    tot_ms = float(np.sum(times))
    if dist is not None:
        t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms = float(t.item())
    value = iters / (tot_ms * 1e-3)

    # ---- end-to-end arm: host buffers -> mcp_ba_load -> compute -> read back ---------------------
    h2 = capi.BaHandle(device=local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h2.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    h2d = sum(np.asarray(a).nbytes for a in (prob.pose_Rt, prob.pose_fixed, prob.pt_xyz, prob.pt_chain, prob.pt_fixed,
                                                prob.meas_xy, prob.meas_chain, prob.meas_pt, prob.meas_noise, prob.meas_cam))
    d2h = prob.pose_Rt.nbytes + prob.pt_xyz.nbytes
    e2e_t, e2e_it = 0.0, 0
    e2e_parts = np.zeros(3)                        # load (marshal + H2D), compute, read-back: host wall clock
    for s in range(warmup + args.steps):
        barrier()
        if s == warmup:
            t_e2e0 = time.perf_counter()
        t = time.perf_counter()
        h2.load(prob)
        t1 = time.perf_counter()
        rc, st = h2.compute(args.lm_iters)
        t2 = time.perf_counter()
        P, X = h2.poses(), h2.points()
        _ = h2.outliers()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if s >= warmup:
            e2e_parts += (t1 - t, t2 - t1, t + dt - t2)
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if s >= warmup:
            e2e_t += dt; e2e_it += rc
    e2e_val = e2e_it / e2e_t
    sampler.window(t_e2e0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing of one profiled step (CUDA events around every launch) ----------------
    h.set_profiling(True)
    h.reset_state()
    rc, st = h.compute(args.lm_iters)
    tm = h.timing()
    h.set_profiling(False)
    n_m, n_p, nc = prob.n_meas, prob.n_pt, 6 * h.n_pose_var
    # algorithmic bytes of one linearise+Schur launch: measurement records (32 B), point records
    # (32 B in, 72 B V/g_p out) and the reduced camera system written once
    lin_bytes = 32.0 * n_m / world + (32.0 + 72.0) * n_p / world + 8.0 * (nc * nc + nc)
    lin_ms = tm["ms_linearize"] / max(tm["n_linearize"], 1)
    peak, peak_src = measured_peaks()
    achieved = lin_bytes / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else 0.0
    iter_bytes = 80.0 * n_m + 232.0 * n_p + 16.0 * nc * nc          # SURVEY.md §8(d), whole LM iteration
    # dram__bytes_read + dram__bytes_write of one launch of each of the two kernels, from the committed ncu --set full
    # capture (cold cache: k_pose_blocks re-reads the 176-byte measurement records k_linearize has just written, which
    # stay in L2 inside a real step)
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_ncu_full_v21_ba.txt")
    if os.path.exists(prof) and args.config == "cfg2":
        unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per_kernel, cur = {}, None
        for ln in open(prof):
            f = ln.split()
            if ln.startswith("Kernel Name"):
                name = "k_linearize" if "k_linearize" in ln else "k_pose_blocks" if "k_pose_blocks" in ln else None
                cur = name if name and name not in per_kernel else None
                if cur:
                    per_kernel[cur] = 0.0
            elif cur and len(f) >= 3 and f[0].startswith("dram__bytes_"):
                per_kernel[cur] += float(f[1]) * unit.get(f[2], 1)
        if len(per_kernel) == 2:
            traffic = sum(per_kernel.values())
    roofline = {"bound": "hbm", "kernel": "k_linearize + k_pose_blocks (reprojection, Jacobians, normal-equation blocks)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": lin_bytes, "kernel_ms": lin_ms,
                "whole_iteration": {"algorithmic_bytes": iter_bytes,
                                    "achieved_gbs": iter_bytes * iters / (tot_ms * 1e-3) / 1e9,
                                    "frac": iter_bytes * iters / (tot_ms * 1e-3) / 1e9 / peak},
                "per_kernel_ms_per_step": {k: v for k, v in tm.items() if k.startswith("ms_")},
                "per_kernel_launches_per_step": {k: v for k, v in tm.items() if k.startswith("n_")}}

    scale = None
    if world > 1 and args.scale_config != "none":
        # BASELINE.json configs[3]: the 1000 KF / 100k-point map, points sharded over the ranks, next to the same map
        # on one GPU (rank 0 alone) measured in the same run
        big = synth.make_ba_config(args.scale_config, seed=args.seed) if rank == 0 else None   # ~25 s of Python once per box,
        dist.barrier()                                                                         # then served from the disk cache
        if big is None:
            big = synth.make_ba_config(args.scale_config, seed=args.seed)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        hb = capi.BaHandle(device=local_rank)
        hb.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        hb.load(big)
        sb = torch.cuda.ExternalStream(hb.stream(), device=torch.device("cuda", local_rank))
        tms, its = [], 0
        for s in range(2 + 3):
            barrier()
            hb.reset_state()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sb)
            rc, stb = hb.compute(args.lm_iters)
            e1.record(sb)
            e1.synchronize()
            if s >= 2:
                tms.append(e0.elapsed_time(e1)); its += rc
        tot = torch.tensor([float(np.sum(tms))], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        hb.close()
        val1 = None
        if rank == 0:
            h1 = capi.BaHandle(device=local_rank)
            h1.load(big)
            s1 = torch.cuda.ExternalStream(h1.stream(), device=torch.device("cuda", local_rank))
            t1, i1 = [], 0
            for s in range(2 + 3):
                h1.reset_state()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s1)
                rc, st1 = h1.compute(args.lm_iters)
                e1.record(s1)
                e1.synchronize()
                if s >= 2:
                    t1.append(e0.elapsed_time(e1)); i1 += rc
            val1 = i1 / (float(np.sum(t1)) * 1e-3)
            h1.close()
        scale = {"workload": "%s: %d poses / %d points / %d measurements, points sharded x%d" % (args.scale_config, big.n_pose, big.n_pt, big.n_meas, world),
                 "value_n_gpus": its / (float(tot.item()) * 1e-3), "value_1_gpu_same_run": val1, "unit": UNIT, "n_gpus": world}
    if world > 1:
        h.close(); h2.close()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        dist = None
    if rank != 0:
        return 0
    frontend = None
    if world == 1 and not args.no_frontend:
        frontend = bench_frontend(capi, synth, local_rank)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        val, ms, n_it = cpu_reference_run(prob, 2, 0, 3)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
               "sample": "2 x 3 LM iterations of the same map on the host, CPU restatement (oracle/ba_oracle.c, -O3 -march=native, Schur solve), "
                         "1 thread; reference binary unavailable"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_pose": prob.n_pose, "n_points": prob.n_pt, "n_meas": prob.n_meas,
                       "lm_iters_per_step": args.lm_iters, "l2": "256 MiB flush between steps; within a step the map stays L2-resident",
                       "parallelism": "points sharded x%d, NCCL allreduce of the Schur system" % world if world > 1 else "1 GPU"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": {"load": 1e3 * e2e_parts[0] / args.steps, "compute": 1e3 * e2e_parts[1] / args.steps,
                                    "read_back": 1e3 * e2e_parts[2] / args.steps}},
            "roofline": roofline, "cpu_baseline": cpu, "frontend": frontend, "scale_big_map": scale, "wall_s": wall,
            "lm": {"iterations_per_step": iters / args.steps, "trials_last_step": st.total_trials}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
