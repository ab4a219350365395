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


FULL_SYSTEM_NOTE = ("the FULL non-marginalised system as the reference configures g2o (BlockSolverX + CHOLMOD, points not marginalised, "
                    "src/ChainBundle.cc:1150-1158,1218): general block-sparse Cholesky with a minimum-degree order on the block graph "
                    "(oracle/ba_oracle.c solve_mode 3; CHOLMOD itself is not in the image), 1 thread")


SCHUR_NOTE = ("the same step with the points eliminated by an explicit per-point Schur complement + dense pose Cholesky (solve_mode 0): "
              "the arithmetic the minimum-degree order performs, without the sparse bookkeeping -- the fastest CPU variant")


def cpu_reference_run(prob, steps, warmup, lm_iters_cpu, solve_mode=0):
    """Times the CPU restatement (oracle) on the host: one thread, as the reference runs BA on the single
    MapMaker thread (no `#pragma omp` in the reference tree).  solve_mode 0: points eliminated by an explicit
    Schur complement; 3: the full system through a general block-sparse Cholesky."""
    from oracle.oracle import OracleBA
    o = OracleBA(prob)
    p0, x0 = np.array(prob.pose_Rt), np.array(prob.pt_xyz)
    total_it, total_t, per_step = 0, 0.0, []
    for s in range(warmup + steps):
        o.set_state(p0, x0)
        t = time.perf_counter()
        rc, st = o.compute(lm_iters_cpu, solve_mode=solve_mode)
        dt = time.perf_counter() - t
        if s >= warmup:
            total_it += max(rc, 0); total_t += dt; per_step.append(dt)
    return total_it / total_t, float(np.mean(per_step)) * 1e3, total_it


def bench_frontend(capi, synth, device, cam_ids=(0, 1, 2, 3), steps=20, warmup=3, n_patches=1000, with_cpu=True):
    """BASELINE.json configs[2]: 640x480 pyramids, FAST-10 + 1k PatchFinder searches per frame per camera, for the cameras
    in `cam_ids` (all four on one GPU; camera c -> GPU c mod N when sharded).  One handle (= one CUDA
    stream) per camera; host buffers in, host results out (H2D/D2H inside the timing)."""
    import time as _t
    cams, frames, reqs = [], [], []
    for c in cam_ids:
        rng = np.random.default_rng(1000 + c)
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
    n_local = len(cams)
    t_kf = t_ps = dev_ps = 0.0
    found, checksum = 0, 0
    for s in range(warmup + steps):
        for c in range(n_local):
            t0 = _t.perf_counter()
            lv = cams[c].make_keyframe(1, frames[c])
            t1 = _t.perf_counter()
            res = cams[c].search_patches(1, reqs[c])
            t2 = _t.perf_counter()
            if s >= warmup:
                t_kf += t1 - t0; t_ps += t2 - t1
                dev_ps += cams[c].timing()["ms_search"]
                found += int(res["found"].sum())
            if s == warmup:                                  # result fingerprint of every camera (compared across GPU counts)
                checksum += int(lv[0]["n_corners"]) * 1000003 + int(res["found"].sum()) * 7919 + int(np.rint(res["found_x"][res["found"] > 0] * 64).sum() % 1000003)
    n_frames = steps * n_local
    tm = cams[0].timing() if cams else {"ms_pyramid": 0.0, "ms_fast": 0.0}
    out = {"cams": list(cam_ids), "n_frames": n_frames, "t_kf": t_kf, "t_ps": t_ps, "dev_ps_ms": dev_ps, "found": found, "checksum": checksum,
           "pyramid_ms_device": tm["ms_pyramid"], "fast_ms_device": tm["ms_fast"], "n_patches": n_patches}
    if with_cpu and cams:
        # CPU restatement of the same per-frame work on one host core (the reference tracker is single threaded)
        from oracle import oracle as ora
        ora.lib()                                            # (first use compiles the -O3 build: not part of the timing)
        ora.pyramid(frames[0])
        t0 = _t.perf_counter()
        pyr_b = ora.pyramid(frames[0])
        lv_b = [ora.level_corners(im) for im in pyr_b]
        t1 = _t.perf_counter()
        pyr_a = ora.pyramid(synth.make_frame(seed=100 + cam_ids[0]))
        ora.search_patches_batch(pyr_a, pyr_b, lv_b, reqs[0][:8])      # warm the marshalling path
        t1b = _t.perf_counter()
        ora.search_patches_batch(pyr_a, pyr_b, lv_b, reqs[0])
        t2 = _t.perf_counter()
        out["cpu_baseline"] = {"keyframe_ms": 1e3 * (t1 - t0), "patches_per_sec": len(reqs[0]) / (t2 - t1b), "cores": 1, "kind": "port",
                               "sample": "1 camera frame (pyramid + FAST-10 + threshold + LUT) and %d patch searches in one C call, oracle/fe_oracle.c" % len(reqs[0])}
    for f in cams:
        f.close()
    return out


def frontend_report(parts, n_gpus, peak):
    """Merges the per-rank results of bench_frontend (every rank timed its own cameras concurrently: the frame rate of
    the rig is the total number of camera frames over the slowest rank's time)."""
    n_frames = sum(p["n_frames"] for p in parts)
    n_patches = parts[0]["n_patches"]
    t_max = max(p["t_kf"] + p["t_ps"] for p in parts)
    t_kf = sum(p["t_kf"] for p in parts); t_ps = sum(p["t_ps"] for p in parts); dev_ps = sum(p["dev_ps_ms"] for p in parts)
    frame_bytes, patch_bytes = 307200 + 100800 + 100800, 800
    kf_dev_ms = parts[0]["pyramid_ms_device"] + parts[0]["fast_ms_device"]
    ps_dev_ms = dev_ps / n_frames
    rep = {"workload": "cfg3: %d-cam 640x480, 4-level pyramid + FAST-10 + %d PatchFinder searches (8 sub-pixel its) per frame, camera c on GPU c mod %d"
                       % (sum(len(p["cams"]) for p in parts), n_patches, n_gpus),
           "n_gpus": n_gpus, "cameras_per_rank": [p["cams"] for p in parts],
           "camera_frames_per_sec_e2e": n_frames / t_max, "keyframe_ms_e2e": 1e3 * t_kf / n_frames,
           "patches_per_sec_e2e": n_frames * n_patches / (t_ps / max(len(parts), 1)) if n_gpus > 1 else n_frames * n_patches / t_ps,
           "patches_per_sec_device": n_frames * n_patches / (dev_ps * 1e-3) * (len(parts) if n_gpus > 1 else 1),
           "patch_search_ms_device": ps_dev_ms, "found_fraction": sum(p["found"] for p in parts) / (n_frames * n_patches),
           "pyramid_ms_device": parts[0]["pyramid_ms_device"], "fast_ms_device": parts[0]["fast_ms_device"],
           "algorithmic_bytes_per_camera_frame": frame_bytes, "patch_bytes_each": patch_bytes,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                        "keyframe": {"kernel": "k_halfsample_fused + k_fast_score_rows + k_fast_compact4", "achieved": frame_bytes / (kf_dev_ms * 1e-3) / 1e9 if kf_dev_ms > 0 else None,
                                     "frac": frame_bytes / (kf_dev_ms * 1e-3) / 1e9 / peak if kf_dev_ms > 0 else None},
                        "patch_search": {"kernel": "k_patch_search_tma", "achieved": patch_bytes * n_patches / (ps_dev_ms * 1e-3) / 1e9 if ps_dev_ms > 0 else None,
                                         "frac": patch_bytes * n_patches / (ps_dev_ms * 1e-3) / 1e9 / peak if ps_dev_ms > 0 else None},
                        "note": "one 640x480 frame is 0.5 MB: launch / dependency latency bound, not HBM bound"},
           "result_checksum_per_camera_set": sum(p["checksum"] for p in parts)}
    for p in parts:
        if "cpu_baseline" in p:
            rep["cpu_baseline"] = p["cpu_baseline"]
            break
    return rep


def bench_stream(capi, synth, device, seconds=2.0):
    """BASELINE.json configs[4]: the 4-camera tracker loop (pyramid + FAST, batched FindPVS projection, 1k patch searches,
    ten pose-update iterations per camera frame, src/Tracker.cc:916-1154) on the main thread while the MapMaker thread keeps
    running bundle adjustments (cfg2, 10 LM iterations each) on its own handle = its own CUDA stream."""
    import threading
    import time as _t
    rng = np.random.default_rng(7)
    prob = synth.make_ba_config("cfg2", seed=0)
    ba = capi.BaHandle(device=device)
    ba.load(prob)
    cam = prob.cams[0]
    cams, frames, reqs, pts = [], [], [], []
    for c in range(4):
        f = capi.FeHandle(640, 480, device=device, max_corners_per_level=16384)
        f.set_camera(cam)
        a = synth.make_frame(seed=300 + c)
        lva = f.make_keyframe(0, a)
        cor = lva[0]["corners"]
        cor = cor[(cor[:, 0] > 16) & (cor[:, 0] < 624) & (cor[:, 1] > 16) & (cor[:, 1] < 464)]
        cor = cor[rng.choice(len(cor), 1000, replace=len(cor) < 1000)]
        rq = np.zeros(len(cor), capi.PATCH_REQ_DTYPE)
        rq["src_kf"] = 0; rq["src_level"] = 0; rq["src_cx"] = cor[:, 0]; rq["src_cy"] = cor[:, 1]
        rq["warp_inv"] = np.array([1.0, 0.0, 0.0, 1.0]); rq["search_level"] = 0
        rq["pred_x"] = cor[:, 0] - 2; rq["pred_y"] = cor[:, 1] + 1
        rq["range"] = 10; rq["subpix_its"] = 8
        # map points in front of the camera for the FindPVS projection and the pose update
        rays = synth.cam_unproject_np(cam, cor.astype(np.float64))
        pts.append(rays * rng.uniform(3.0, 9.0, (len(rays), 1)))
        cams.append(f); reqs.append(rq)
        frames.append([synth.make_frame(seed=300 + c, shift=(2.0 + 0.5 * k, -1.0)) for k in range(3)])
    ident = np.concatenate([np.eye(3).reshape(-1), np.zeros(3)])
    stop = {"flag": False, "n": 0, "iters": 0}

    def mapmaker():
        while not stop["flag"]:
            ba.reset_state()
            rc, st = ba.compute(10)
            stop["n"] += 1; stop["iters"] += max(rc, 0)

    def track_one(k):
        for c, f in enumerate(cams):
            f.make_keyframe(1, frames[c][k % 3])
            zeros = np.zeros_like(pts[c])
            f.project_points(ident, pts[c], zeros + np.array([1e-3, 0, 0]), zeros + np.array([0, 1e-3, 0]))     # FindPVS block
            res = f.search_patches(1, reqs[c])
            jr = f.calc_jacobians(ident, ident, pts[c])                                                    # TrackerData::CalcJacobian
            meas = np.zeros(len(pts[c]), capi.POSE_MEAS_DTYPE)
            meas["image"] = jr["px"]; meas["jac"] = jr["jac"]; meas["sqrt_inv_noise"] = 1.0
            meas["found"][:, 0] = jr["px"][:, 0] + (res["found_x"] - res["coarse_x"]) * 0.1
            meas["found"][:, 1] = jr["px"][:, 1] + (res["found_y"] - res["coarse_y"]) * 0.1
            meas["found_flag"] = (res["found"] > 0) & (jr["in_image"] == 1)
            for _ in range(10):                                                                            # src/Tracker.cc:1089-1140
                f.pose_update(meas)

    for k in range(3):
        track_one(k)
    t0 = _t.perf_counter(); n_alone = 0
    while _t.perf_counter() - t0 < 0.5 * seconds:
        track_one(n_alone); n_alone += 1
    fps_alone = n_alone / (_t.perf_counter() - t0)
    th = threading.Thread(target=mapmaker)
    th.start()
    t0 = _t.perf_counter(); n = 0
    while _t.perf_counter() - t0 < seconds:
        track_one(n); n += 1
    dt = _t.perf_counter() - t0
    stop["flag"] = True
    th.join()
    for f in cams:
        f.close()
    ba.close()
    return {"workload": "cfg5: 4-cam 640x480 tracker loop (pyramid + FAST-10, FindPVS projection of 1k points, 1k patch searches, 10 pose updates "
                        "per camera frame) with cfg2 bundle adjustments (10 LM iterations each) running on the BA handle's stream",
            "rig_frames_per_sec_with_ba": n / dt, "rig_frames_per_sec_alone": fps_alone, "target_fps": 30.0,
            "ba_calls_meanwhile": stop["n"], "ba_lm_iters_per_sec_meanwhile": stop["iters"] / dt}


def bench_config(workload, prob, lm_iters, world):
    """`config` of the JSON line: identical for both arms except for `parallelism` when ranks > 1."""
    return {"workload": workload, "n_pose": prob.n_pose, "n_points": prob.n_pt, "n_meas": prob.n_meas, "lm_iters_per_step": lm_iters,
            "l2": "256 MiB flush between steps; within a step the map stays L2-resident",
            "parallelism": "points sharded x%d, NCCL allreduce of the Schur system" % world if world > 1 else "1 GPU"}


def csrc_digest(prefixes=("ba_", "mcp_common"), host_only=("ba_prep.hpp", "ba_prep_cpu.cpp")):
    """Content hash of the bundle-adjuster kernel sources and their launcher: a committed ncu capture is only quoted if it
    was taken from them.  The host marshalling (ba_prep.*: no device code, no launch parameter) is not part of it."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "mcptam_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith(prefixes) and f not in host_only:
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the named kernel from profiles/r02_ncu_full_ba.txt
    (tools/ncu_extract.py output, first line '# csrc <digest>').  None (and why) if the capture is from other sources."""
    prof = os.path.join(ROOT, "profiles", "r02_ncu_full_ba.txt")
    if not os.path.exists(prof):
        return None, "no capture committed"
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    lines = open(prof).read().splitlines()
    tag = lines[0].split()[2] if lines and lines[0].startswith("# csrc") and len(lines[0].split()) > 2 else None
    if tag != csrc_digest():
        return None, "capture is from other kernel sources (%s, now %s)" % (tag, csrc_digest())
    tot, cur, seen = 0.0, False, False
    for ln in lines:
        f = ln.split()
        if ln.startswith("Kernel Name"):
            if seen and cur:
                break
            cur = kernel_substr in ln
            seen = seen or cur
        elif cur and len(f) >= 3 and f[0].startswith("dram__bytes_") and f[0].endswith(".sum"):
            tot += float(f[1]) * unit.get(f[2], 1)
    return (tot if seen else None), ("ok" if seen else "kernel not in the capture")


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
    args.warmup = warmup

    from mcptam_b200 import synth
    prob = synth.make_ba_config(args.config, seed=args.seed)
    workload = WORKLOAD if args.config == "cfg2" else args.config

    if args.impl == "reference":
        if rank != 0:
            return 0
        ncores = os.cpu_count()
        # the same step as the GPU arm (--lm-iters outer iterations from the same initial estimate); one step is ~2 s of
        # CPU work at cfg2, so the driver's K and W stay as they are
        val, ms, n_it = cpu_reference_run(prob, args.steps, args.warmup, args.lm_iters, solve_mode=3)
        val_schur, _, _ = cpu_reference_run(prob, 1, 0, args.lm_iters, solve_mode=0)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": bench_config(workload, prob, args.lm_iters, args.gpus),
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores": ncores, "kind": "port",
                                 "sample": "%d steps x %d LM iterations of the same map, CPU restatement (oracle/ba_oracle.c, "
                                           "-O3 -march=native), %s like the reference's MapMaker thread; reference binary "
                                           "unavailable (no ROS/TooN/g2o/SuiteSparse)" % (args.steps, args.lm_iters, FULL_SYSTEM_NOTE),
                                 "schur_restatement": {"value": val_schur, "unit": UNIT, "sample": "1 step x %d LM iterations; %s" % (args.lm_iters, SCHUR_NOTE)}},
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
    # The synthetic map lists its measurements point by point (no sort in mcp_ba_load).  The reference's own loop adds them
    # keyframe by keyframe (src/BundleAdjusterMulti.cc:168-199): the same end-to-end call in THAT order, a few steps.
    e2e_kf = None
    if world == 1:
        import copy as _copy
        pk = _copy.copy(prob)
        order = np.argsort(np.asarray(prob.meas_chain)[:, 0].astype(np.int64) * 64 + np.asarray(prob.meas_cam), kind="stable")
        for k in ("meas_xy", "meas_chain", "meas_pt", "meas_noise", "meas_cam"):
            setattr(pk, k, np.ascontiguousarray(np.asarray(getattr(prob, k))[order]))
        kt, ki, kl = 0.0, 0, 0.0
        for s in range(3 + 10):
            torch.cuda.synchronize()
            t = time.perf_counter()
            h2.load(pk)
            t1 = time.perf_counter()
            rc, st = h2.compute(args.lm_iters)
            P, X = h2.poses(), h2.points()
            _ = h2.outliers()
            torch.cuda.synchronize()
            if s >= 3:
                kt += time.perf_counter() - t; ki += rc; kl += t1 - t
        e2e_kf = {"value": ki / kt, "unit": UNIT, "load_ms": 1e3 * kl / 10,
                  "note": "measurements in the order the reference's BundleAdjusterMulti loop adds them (keyframe-major): mcp_ba_load runs its parallel sort by point"}
    sampler.window(t_e2e0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing of one profiled step (CUDA events around every launch, one LM candidate per round) ----
    h.set_profiling(True)
    h.reset_state()
    rc, st = h.compute(args.lm_iters)
    tm = h.timing()
    h.set_profiling(False)
    n_m, n_p, nc = prob.n_meas, prob.n_pt, 6 * h.n_pose_var
    # (point, movable pose) slots: distinct movable keyframes (observers + source) a point's measurements touch
    mov = ~prob.pose_fixed.astype(bool)
    pairs = np.concatenate([np.stack([prob.meas_pt, prob.meas_chain[:, 0]], 1)[mov[prob.meas_chain[:, 0]]],
                            np.stack([np.arange(n_p), prob.pt_chain[:, 0]], 1)[mov[prob.pt_chain[:, 0]]]])
    n_slots = len(np.unique(pairs[:, 0].astype(np.int64) * prob.n_pose + pairs[:, 1]))
    peak, peak_src = measured_peaks()
    # algorithmic bytes per launch (per-unit figures x the units one launch processes on this rank)
    alg = {"linearize": ("k_linearize + k_pose_blocks", 32.0 * n_m / world + 104.0 * n_p / world + 8.0 * (nc * nc + nc)),
           "schur": ("k_schur_vinv_multi + k_schur_pairs (one candidate)", (144.0 * n_slots + 96.0 * n_p) / world + 8.0 * (nc * nc + nc)),
           "solve": ("k_chol_solve", 2 * 8.0 * (nc * nc + nc)),
           "backsub": ("k_backsub_eval", (40.0 * n_m + 80.0 * n_p + 144.0 * n_slots) / world),
           "select": ("k_select_cluster / k_select_grid", 8.0 * n_m)}
    per_kernel = {}
    for key, (name, nbytes) in alg.items():
        n_l, ms = tm["n_" + key], tm["ms_" + key]
        if n_l:
            t = ms / n_l
            per_kernel[key] = {"kernel": name, "launches_per_step": n_l, "ms_per_launch": t, "share_of_kernel_time": None,
                               "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": nbytes / (t * 1e-3) / 1e9,
                               "frac": nbytes / (t * 1e-3) / 1e9 / peak}
    tot_k = sum(tm["ms_" + k] for k in ("select", "linearize", "schur", "solve", "backsub", "control", "other"))
    for key in per_kernel:
        per_kernel[key]["share_of_kernel_time"] = tm["ms_" + key] / tot_k
    dom = max(per_kernel, key=lambda k: tm["ms_" + k])            # the dominant kernel by time
    traffic, traffic_note = ncu_traffic(alg[dom][0].split()[0]) if args.config == "cfg2" else (None, "capture is of cfg2")
    iter_bytes = 80.0 * n_m + 232.0 * n_p + 16.0 * nc * nc          # SURVEY.md §8(d), whole LM iteration
    roofline = {"bound": "hbm", "kernel": per_kernel[dom]["kernel"] + " (dominant kernel by time: %.0f %% of the summed kernel time of a step)"
                % (100 * per_kernel[dom]["share_of_kernel_time"]),
                "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kernel[dom]["frac"],
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_kernel[dom]["algorithmic_bytes_per_launch"], "kernel_ms": per_kernel[dom]["ms_per_launch"],
                "note": "the working set (%.1f MB) is L2-resident and every kernel is bound by dependency latency / fp64 issue, not by HBM: "
                        "fractions of the HBM roofline are small by construction (DESIGN.md §5)" % (iter_bytes / 1e6),
                "whole_iteration": {"algorithmic_bytes": iter_bytes,
                                    "achieved_gbs": iter_bytes * iters / (tot_ms * 1e-3) / 1e9,
                                    "frac": iter_bytes * iters / (tot_ms * 1e-3) / 1e9 / peak},
                "per_kernel": per_kernel,
                "per_kernel_ms_per_step": {k: v for k, v in tm.items() if k.startswith("ms_")},
                "per_kernel_launches_per_step": {k: v for k, v in tm.items() if k.startswith("n_")}}

    def one_gpu_reference(problem, P, X, trials):
        """rank 0: the same map on ONE GPU; the sharded result must agree with it (sum order differs: 1e-7 relative)."""
        h1 = capi.BaHandle(device=local_rank)
        h1.load(problem)
        s1 = torch.cuda.ExternalStream(h1.stream(), device=torch.device("cuda", local_rank))
        t1, i1 = [], 0
        for k in range(2 + 3):
            h1.reset_state()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s1)
            rc1, st1 = h1.compute(args.lm_iters)
            e1.record(s1)
            e1.synchronize()
            if k >= 2:
                t1.append(e0.elapsed_time(e1)); i1 += rc1
        P1, X1 = h1.poses(), h1.points()
        h1.close()
        relp = float(np.linalg.norm(P - P1) / np.linalg.norm(P1)); relx = float(np.linalg.norm(X - X1) / np.linalg.norm(X1))
        ok = bool(relp < 1e-7 and relx < 1e-7 and trials == st1.total_trials)
        return i1 / (float(np.sum(t1)) * 1e-3), {"rel_pose": relp, "rel_point": relx, "trials": [int(trials), int(st1.total_trials)], "ok": ok}

    multi_parity = None
    if world > 1:
        h.reset_state()
        rc, st = h.compute(args.lm_iters)
        Pm, Xm = h.poses(), h.points()
        if rank == 0:
            _, multi_parity = one_gpu_reference(prob, Pm, Xm, st.total_trials)
        dist.barrier()

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
        Pb, Xb = hb.poses(), hb.points()
        hb.close()
        val1, big_parity = None, None
        if rank == 0:
            val1, big_parity = one_gpu_reference(big, Pb, Xb, stb.total_trials)
        scale = {"workload": "%s: %d poses / %d points / %d measurements, points sharded x%d" % (args.scale_config, big.n_pose, big.n_pt, big.n_meas, world),
                 "value_n_gpus": its / (float(tot.item()) * 1e-3), "value_1_gpu_same_run": val1, "unit": UNIT, "n_gpus": world,
                 "multi_gpu_parity": big_parity}
    # ---- front end (configs[2]); sharded one camera per GPU when there are several ranks -------
    frontend = None
    if not args.no_frontend:
        mine = [c for c in range(4) if c % world == rank]
        part = bench_frontend(capi, synth, local_rank, cam_ids=mine, with_cpu=(rank == 0 and world == 1))
        parts = [part]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, part)
        if rank == 0:
            frontend = frontend_report([p for p in parts if p["n_frames"]], world, peak)
            if world > 1:
                # the sharded rig must produce what one GPU produces for the same cameras
                one = bench_frontend(capi, synth, local_rank, cam_ids=(0, 1, 2, 3), steps=2, warmup=1, with_cpu=False)
                frontend["multi_gpu_parity"] = {"checksum_sharded": frontend["result_checksum_per_camera_set"], "checksum_one_gpu": one["checksum"],
                                                "ok": bool(one["checksum"] == frontend["result_checksum_per_camera_set"])}
    if world > 1:
        h.close(); h2.close()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        dist = None
    if rank != 0:
        return 0
    stream_section = None
    if world == 1 and not args.no_frontend:
        stream_section = bench_stream(capi, synth, local_rank)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        val, ms, n_it = cpu_reference_run(prob, 1, 0, args.lm_iters, solve_mode=3)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
               "sample": "1 step (%d LM iterations) of the same map on the host, CPU restatement (oracle/ba_oracle.c, -O3 -march=native), %s; "
                         "reference binary unavailable" % (args.lm_iters, FULL_SYSTEM_NOTE)}
        val_schur, _, _ = cpu_reference_run(prob, 1, 0, args.lm_iters, solve_mode=0)
        cpu["schur_restatement"] = {"value": val_schur, "unit": UNIT, "sample": "1 step x %d LM iterations; %s" % (args.lm_iters, SCHUR_NOTE)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(workload, prob, args.lm_iters, world),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": {"load": 1e3 * e2e_parts[0] / args.steps, "compute": 1e3 * e2e_parts[1] / args.steps,
                                    "read_back": 1e3 * e2e_parts[2] / args.steps},
                    "reference_call_order": e2e_kf},
            "roofline": roofline, "cpu_baseline": cpu, "multi_gpu_parity": multi_parity, "frontend": frontend, "tracker_mapmaker_stream": stream_section,
            "scale_big_map": scale, "wall_s": wall,
            "lm": {"iterations_per_step": iters / args.steps, "trials_last_step": st.total_trials}}
    print(json.dumps(line))
    bad = [p for p in (multi_parity, scale and scale.get("multi_gpu_parity"), frontend and frontend.get("multi_gpu_parity")) if p and not p["ok"]]
    if bad:
        sys.stderr.write("bench.py: a sharded result differs from the 1-GPU result: %r\n" % (bad,))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
