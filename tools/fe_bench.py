"""Device / end-to-end timing of the per-frame front end for one 640x480 camera (torch-free).
usage: fe_bench.py [reps]   (MCP_FE_FAST_FUSED / MCP_FE_TMA select the kernel variants)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcptam_b200 import synth, capi

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(1000)
f = capi.FeHandle(640, 480, max_corners_per_level=16384)
a = synth.make_frame(seed=100)
b = synth.make_frame(seed=100, shift=(3.0, -2.0))
lva = f.make_keyframe(0, a)
cor = lva[0]["corners"]
cor = cor[(cor[:, 0] > 16) & (cor[:, 0] < 624) & (cor[:, 1] > 16) & (cor[:, 1] < 464)]
cor = cor[rng.choice(len(cor), 1000, replace=len(cor) < 1000)]
rq = np.zeros(1000, capi.PATCH_REQ_DTYPE)
rq["src_kf"] = 0; rq["src_level"] = 0; rq["src_cx"] = cor[:, 0]; rq["src_cy"] = cor[:, 1]
rq["warp_inv"] = np.array([1.0, 0.02, -0.02, 1.0]); rq["search_level"] = 0
rq["pred_x"] = cor[:, 0] - 3 + rng.integers(-2, 3, 1000); rq["pred_y"] = cor[:, 1] + 2 + rng.integers(-2, 3, 1000)
rq["range"] = 10; rq["subpix_its"] = 8
acc = {"ms_pyramid": 0.0, "ms_fast": 0.0, "ms_other": 0.0, "ms_compact": 0.0, "ms_search": 0.0}
t_kf = t_ps = 0.0
for s in range(10 + reps):
    t0 = time.perf_counter(); lv = f.make_keyframe(1, b); t1 = time.perf_counter()
    tm = f.timing()
    res = f.search_patches(1, rq); t2 = time.perf_counter()
    tm2 = f.timing()
    if s >= 10:
        t_kf += t1 - t0; t_ps += t2 - t1
        for k in ("ms_pyramid", "ms_fast", "ms_other", "ms_compact"): acc[k] += tm[k]
        acc["ms_search"] += tm2["ms_search"]
print("corners", [int(l["n_corners"]) for l in lv], "thr", [int(l["fast_thresh"]) for l in lv], "found", int(res["found"].sum()))
print("keyframe e2e ms", 1e3 * t_kf / reps, "search e2e ms", 1e3 * t_ps / reps, {k: round(1e3 * v / reps, 2) for k, v in acc.items()}, "(device, us)")
