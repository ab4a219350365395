"""Torch-free driver for ncu captures: loads a synthetic map and runs `reps` BundleAdjust-equivalent calls.
usage: prof_ba.py [cfg] [lm_iters] [reps] [fe]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcptam_b200 import synth, capi

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prob = synth.make_ba_config(cfg, 0)
g = capi.BaHandle()
g.load(prob)
import time
wall = []
for _ in range(reps):
    g.reset_state()
    t0 = time.perf_counter()
    rc, st = g.compute(iters)
    wall.append(time.perf_counter() - t0)
print("rc", rc, "trials", st.total_trials, "launches", st.kernel_launches, "gpu_ms", st.gpu_ms, "call_ms(min)", 1e3 * min(wall),
      "n_outliers", st.n_outliers, "chi2", st.chi2_before, st.chi2_after)
if len(sys.argv) > 4 and sys.argv[4] == "fe":
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    a = synth.make_frame(seed=100); b = synth.make_frame(seed=100, shift=(3.0, -2.0))
    lva = f.make_keyframe(0, a)
    cor = lva[0]["corners"]
    cor = cor[(cor[:, 0] > 16) & (cor[:, 0] < 624) & (cor[:, 1] > 16) & (cor[:, 1] < 464)][:1000]
    rq = np.zeros(len(cor), capi.PATCH_REQ_DTYPE)
    rq["src_kf"] = 0; rq["src_level"] = 0; rq["src_cx"] = cor[:, 0]; rq["src_cy"] = cor[:, 1]
    rq["warp_inv"] = np.array([1.0, 0.02, -0.02, 1.0]); rq["search_level"] = 0
    rq["pred_x"] = cor[:, 0] - 3; rq["pred_y"] = cor[:, 1] + 2
    rq["range"] = 10; rq["subpix_its"] = 8
    for _ in range(2):
        f.make_keyframe(1, b)
        res = f.search_patches(1, rq)
    print("fe found", int(res["found"].sum()), "of", len(rq))
