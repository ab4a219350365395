"""BASELINE.json configs[4]: the per-frame front end of a 4-camera rig keeps running on its own CUDA streams while
a bundle adjustment runs on the BA handle's side stream (the reference runs them on the Tracker and MapMaker
threads, src/System.cc:169-170).  Results of both must be identical to running them alone."""
import threading
import time

import numpy as np
import pytest

from mcptam_b200 import synth

pytestmark = pytest.mark.gpu


def _track_frames(capi, cams, frames, reqs, n_frames):
    out = []
    for k in range(n_frames):
        for c, f in enumerate(cams):
            lv = f.make_keyframe(1, frames[c][k % len(frames[c])])
            res = f.search_patches(1, reqs[c])
            out.append((lv[0]["n_corners"], lv[0]["corners"].copy(), res.copy()))
    return out


def test_frontend_concurrent_with_bundle_adjustment():
    from mcptam_b200 import capi
    rng = np.random.default_rng(0)
    n_cams, n_frames = 4, 12
    cams, frames, reqs = [], [], []
    for c in range(n_cams):
        f = capi.FeHandle(640, 480, max_corners_per_level=16384)
        a = synth.make_frame(seed=200 + c)
        lva = f.make_keyframe(0, a)
        cor = lva[0]["corners"]
        cor = cor[(cor[:, 0] > 16) & (cor[:, 0] < 624) & (cor[:, 1] > 16) & (cor[:, 1] < 464)]
        cor = cor[rng.choice(len(cor), 500, replace=False)]
        rq = np.zeros(len(cor), capi.PATCH_REQ_DTYPE)
        rq["src_kf"] = 0; rq["src_level"] = 0; rq["src_cx"] = cor[:, 0]; rq["src_cy"] = cor[:, 1]
        rq["warp_inv"] = np.array([1.0, 0.0, 0.0, 1.0]); rq["search_level"] = 0
        rq["pred_x"] = cor[:, 0] - 2; rq["pred_y"] = cor[:, 1] + 1
        rq["range"] = 10; rq["subpix_its"] = 8
        cams.append(f); reqs.append(rq)
        frames.append([synth.make_frame(seed=200 + c, shift=(2.0 + 0.5 * k, -1.0)) for k in range(3)])
    prob = synth.make_ba_config("cfg1", seed=0)
    ba = capi.BaHandle()
    ba.load(prob)
    rc0, st0 = ba.compute(10)
    P0, X0 = ba.poses(), ba.points()
    alone = _track_frames(capi, cams, frames, reqs, n_frames)

    result = {}

    def mapmaker():
        n = 0
        t = time.perf_counter()
        while not result.get("stop"):
            ba.reset_state()
            rc, st = ba.compute(10)
            n += 1
        result.update(rc=rc, trials=st.total_trials, P=ba.poses(), X=ba.points(), n=n, dt=time.perf_counter() - t)

    th = threading.Thread(target=mapmaker)
    th.start()
    t = time.perf_counter()
    both = _track_frames(capi, cams, frames, reqs, n_frames)
    dt = time.perf_counter() - t
    result["stop"] = True
    th.join()
    assert result["n"] >= 1
    for (n0, c0, r0), (n1, c1, r1) in zip(alone, both):
        assert n0 == n1 and np.array_equal(c0, c1)
        assert np.array_equal(r0["found"], r1["found"]) and np.array_equal(r0["found_x"], r1["found_x"]) and np.array_equal(r0["found_y"], r1["found_y"])
    assert result["rc"] == rc0 and result["trials"] == st0.total_trials
    # fp64 atomics make the normal-equation sums order dependent: identical up to rounding, not bitwise
    assert np.linalg.norm(result["P"] - P0) <= 1e-9 * np.linalg.norm(P0) and np.linalg.norm(result["X"] - X0) <= 1e-9 * np.linalg.norm(X0)
    fps = n_frames / dt
    print("4-camera frames/s with BA running: %.0f (BA calls completed meanwhile: %d)" % (fps, result["n"]))
    assert fps > 30.0          # the rig's 30 fps stream leaves headroom
