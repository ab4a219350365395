"""GPU parity tests of the front end: CUDA path (through the C ABI) vs the CPU oracle (oracle/fe_oracle.c).

Integer / byte / index work is compared bit-exactly; the mixed fp32/fp64 sub-pixel refinement is compared
bit-exactly too (the kernel forbids FMA contraction on that path, the oracle is built with -ffp-contract=off).
The comparator is a RESTATEMENT of the reference (pinned against the reference's own PatchFinder / MiniPatch / ShiTomasi in tests/test_oracle_vs_ref.py; libCVD [3P]).
"""
import numpy as np
import pytest

from mcptam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from mcptam_b200 import capi
    capi.lib()
    return capi


@pytest.fixture(scope="module")
def ora():
    from oracle import oracle
    oracle.lib()
    return oracle


def _check_levels(ora, img, lv, mask=None, adaptive=True):
    pyr = ora.pyramid(img)
    mpyr = ora.pyramid(mask) if mask is not None else [None] * 4
    fixed = [10, 15, 15, 10]
    for l in range(4):
        ref = ora.level_corners(pyr[l], mask=mpyr[l] if adaptive else None, adaptive=adaptive, fixed_thresh=fixed[l])
        assert lv[l]["n_corners"] == ref["n_corners"], (l, lv[l]["n_corners"], ref["n_corners"])
        assert lv[l]["fast_thresh"] == ref["fast_thresh"]
        assert np.array_equal(lv[l]["fast_freq"], ref["fast_freq"])
        assert np.array_equal(lv[l]["corners"], ref["corners"])                 # raster order, index-exact
        assert np.array_equal(lv[l]["row_lut"][: pyr[l].shape[0]], ref["row_lut"])
    return pyr


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_pyramid_and_corners(capi, ora, seed):
    img = synth.make_frame(seed=seed)
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    lv = f.make_keyframe(0, img, want_images=True)
    pyr = _check_levels(ora, img, lv)
    for l in range(4):
        h, w = pyr[l].shape
        assert np.array_equal(lv[l]["image"].reshape(-1)[: h * w].reshape(h, w), pyr[l])


def test_corners_with_mask_and_fixed_threshold(capi, ora):
    img = synth.make_frame(seed=4)
    mask = np.full((480, 640), 255, np.uint8)
    mask[100:300, 200:420] = 0
    mask[::7, ::5] = 128
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    f.set_mask(mask)
    _check_levels(ora, img, f.make_keyframe(0, img), mask=mask)
    f.set_mask(None)
    _check_levels(ora, img, f.make_keyframe(1, img))
    g = capi.FeHandle(640, 480, adaptive_thresh=0, max_corners_per_level=16384)
    _check_levels(ora, img, g.make_keyframe(0, img), adaptive=False)


def test_odd_size_and_edge_images(capi, ora):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (250, 330), dtype=np.uint8)       # not a multiple of 8: generic pyramid path
    f = capi.FeHandle(330, 250, max_corners_per_level=1 << 16)
    lv = f.make_keyframe(0, img)
    _check_levels(ora, img, lv)
    flat = np.full((250, 330), 77, np.uint8)                     # no corners at all
    lv = f.make_keyframe(1, flat)
    assert all(l["n_corners"] == 0 for l in lv)
    sat = np.zeros((250, 330), np.uint8); sat[::2, ::2] = 255    # extreme contrast, maximum scores
    _check_levels(ora, sat, f.make_keyframe(2, sat))


@pytest.mark.parametrize("kind", ["frame", "noise", "saturated"])
def test_score_map_exact(capi, ora, kind):
    """every pixel of the FAST score map (0 = no corner at b = 5, else libCVD's fast_corner_score_10) against the oracle's
    bisection: the kernel evaluates the score in closed form with 16-bit SIMD min3/max3"""
    if kind == "frame":
        img = synth.make_frame(seed=7)
    elif kind == "noise":
        img = np.random.default_rng(3).integers(0, 256, (480, 640), dtype=np.uint8)
    else:
        img = np.zeros((480, 640), np.uint8); img[::2, ::2] = 255; img[1::3, 1::2] = 128
    f = capi.FeHandle(640, 480, max_corners_per_level=1 << 17)
    f.make_keyframe(0, img)
    pyr = ora.pyramid(img)
    for l in range(4):
        sm = f.debug_scores(0, l)
        xy = ora.fast10_detect(pyr[l], 5)
        ref = np.zeros_like(sm)
        if len(xy):
            ref[xy[:, 1], xy[:, 0]] = np.minimum(ora.fast10_score(pyr[l], xy, 5), 255)
        assert np.array_equal(sm, ref), (kind, l, int((sm != ref).sum()))


def _make_requests(capi, rng, lv_src, n, shift, src_slot=0, subpix=True, exhaustive_frac=0.05):
    req = np.zeros(n, capi.PATCH_REQ_DTYPE)
    k = 0
    while k < n:
        sl = int(rng.integers(0, 4))
        cor = lv_src[sl]["corners"]
        if len(cor) == 0:
            continue
        cx, cy = cor[rng.integers(len(cor))]
        w, h = lv_src[sl]["width"], lv_src[sl]["height"]
        if not (10 <= cx < w - 10 and 10 <= cy < h - 10):
            if rng.random() < 0.9:
                continue
        ang = rng.normal(0, 0.08)
        sc = np.exp(rng.normal(0, 0.08)) * (1 << sl)     # warp_inv maps source-level pixels to level-0 target pixels
        A = sc * np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        det = np.linalg.det(A)
        lvl = 0
        while det > 3 and lvl < 3:
            lvl += 1
            det *= 0.25
        l0 = ((cx + 0.5) * (1 << sl) - 0.5 - shift[0], (cy + 0.5) * (1 << sl) - 0.5 - shift[1])
        r = req[k]
        r["src_kf"] = src_slot; r["src_level"] = sl; r["src_cx"] = cx; r["src_cy"] = cy
        r["warp_inv"] = A.reshape(-1)
        r["search_level"] = lvl
        r["pred_x"] = int(l0[0] + rng.uniform(-3, 3)); r["pred_y"] = int(l0[1] + rng.uniform(-3, 3))
        r["range"] = int(rng.choice([5, 10, 30]))
        r["subpix_its"] = int(rng.choice([0, 8, 10])) if subpix else 0
        r["exhaustive"] = int(rng.random() < exhaustive_frac)
        if r["exhaustive"]:
            r["range"] = 10
            r["subpix_its"] = 10
        k += 1
    return req


@pytest.mark.parametrize("seed", [11, 12])
def test_patch_search_parity(capi, ora, seed):
    rng = np.random.default_rng(seed)
    shift = (3.0, -2.0)
    a = synth.make_frame(seed=seed)
    b = synth.make_frame(seed=seed, shift=shift)
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    lva = f.make_keyframe(0, a)
    lvb = f.make_keyframe(1, b)
    req = _make_requests(capi, rng, lva, 600, shift)
    # some requests off the image / degenerate
    req[0]["pred_x"] = -50; req[1]["pred_y"] = 5000; req[2]["src_cx"] = 1; req[2]["src_cy"] = 1
    res = f.search_patches(1, req)
    templ = f.templates(len(req))
    pa, pb = ora.pyramid(a), ora.pyramid(b)
    tl = [ora.level_corners(im) for im in pb]
    n_found = n_sub = 0
    for i in range(len(req)):
        ref = ora.search_patch(pa, pb, tl, req[i])
        assert res[i]["template_bad"] == ref["template_bad"], i
        if not ref["template_bad"]:
            assert np.array_equal(templ[i], ref["template"]), i           # CVD::transform bytes
        assert res[i]["found"] == ref["found"], (i, res[i], ref)
        if not ref["template_bad"]:
            assert res[i]["score"] == ref["score"], (i, res[i]["score"], ref["score"])
        if ref["found"]:
            n_found += 1
            assert (res[i]["coarse_x"], res[i]["coarse_y"]) == (ref["coarse_x"], ref["coarse_y"])
            assert res[i]["did_subpix"] == ref["did_subpix"]
            n_sub += ref["did_subpix"]
            assert res[i]["found_x"] == ref["found_x"] and res[i]["found_y"] == ref["found_y"], (i, res[i], ref)
    assert n_found > 200 and n_sub > 100                                   # the test exercises the interesting paths


def test_patch_search_identity_recovers_shift(capi):
    """Size-independent property: with an identity warp the search finds each corner at the shifted position."""
    rng = np.random.default_rng(5)
    shift = (4.0, 1.0)
    a = synth.make_frame(seed=21)
    b = synth.make_frame(seed=21, shift=shift)
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    lva = f.make_keyframe(0, a)
    f.make_keyframe(1, b, outputs=False)
    cor = lva[0]["corners"]
    cor = cor[(cor[:, 0] > 20) & (cor[:, 0] < 620) & (cor[:, 1] > 20) & (cor[:, 1] < 460)]
    cor = cor[rng.choice(len(cor), 1000, replace=False)]
    req = np.zeros(len(cor), capi.PATCH_REQ_DTYPE)
    req["src_kf"] = 0; req["src_level"] = 0; req["src_cx"] = cor[:, 0]; req["src_cy"] = cor[:, 1]
    req["warp_inv"] = np.array([1.0, 0, 0, 1.0]); req["search_level"] = 0
    req["pred_x"] = cor[:, 0] - 4; req["pred_y"] = cor[:, 1] - 1
    req["range"] = 10; req["subpix_its"] = 8
    res = f.search_patches(1, req)
    ok = res["found"] == 1
    assert ok.mean() > 0.9
    err = np.hypot(res["found_x"][ok] - (cor[ok, 0] - shift[0]), res["found_y"][ok] - (cor[ok, 1] - shift[1]))
    assert np.median(err) < 0.2


def test_shitomasi_and_minipatch(capi, ora):
    rng = np.random.default_rng(9)
    a = synth.make_frame(seed=31)
    b = synth.make_frame(seed=31, shift=(2.0, 1.0))
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    lva = f.make_keyframe(0, a)
    lvb = f.make_keyframe(1, b)
    pa, pb = ora.pyramid(a), ora.pyramid(b)
    for l in (0, 2):
        cor = lva[l]["corners"]
        h, w = pa[l].shape
        cor = cor[(cor[:, 0] >= 10) & (cor[:, 0] < w - 10) & (cor[:, 1] >= 10) & (cor[:, 1] < h - 10)][:300]
        st = f.shitomasi(0, l, cor)
        ref = np.array([ora.shitomasi(pa[l], x, y) for x, y in cor])
        assert np.array_equal(st, ref)
        start = cor + rng.integers(-3, 4, cor.shape)
        pos, found = f.minipatch_find(0, 1, l, cor, start, 10)
        for i in range(len(cor)):
            fr, pr = ora.minipatch_find(pa[l], pb[l], lvb[l]["corners"], lvb[l]["row_lut"][: pb[l].shape[0]], cor[i], start[i], 10)
            assert bool(found[i]) == fr, i
            if fr:
                assert tuple(pos[i]) == tuple(pr), i


def test_project_points_parity(capi, ora):
    """FindPVS building block: Project + GetDerivsUnsafe + CalcSearchLevelAndWarpMatrix vs the oracle (1e-12 relative)."""
    rng = np.random.default_rng(3)
    cams, extr = synth.make_rig(2, rng)
    cam = cams[0]
    f = capi.FeHandle(640, 480)
    f.set_camera(cam)
    R = synth.so3_exp(rng.normal(0, 0.3, 3)); t = rng.normal(0, 0.5, 3)
    T = synth.rt_pack((R, t))
    n = 4000
    pc = rng.normal(0, 1, (n, 3)) * [4, 4, 2] + [0, 0, 6]              # mostly in front of the camera
    pw = (pc - t) @ R                                                   # world = R^T (pc - t)
    rw = rng.normal(0, 1, (n, 3)) * 0.01 * np.linalg.norm(pc, axis=1)[:, None]
    dw = rng.normal(0, 1, (n, 3)) * 0.01 * np.linalg.norm(pc, axis=1)[:, None]
    res = f.project_points(T, pw, rw, dw)
    n_lvl = 0
    for i in range(n):
        ref = ora.project_point(cam, T, pw[i], rw[i], dw[i])
        assert res[i]["in_image"] == ref["in_image"], i
        assert res[i]["search_level"] == ref["level"], (i, res[i]["search_level"], ref["level"], res[i]["warp_inv"], ref["warp_inv"])
        n_lvl += ref["level"] >= 0
        for key in ("px", "cam_derivs", "warp_inv", "v3cam"):
            a, b = res[i][key], ref[key if key != "cam_derivs" else "derivs"]
            assert np.allclose(a, b, rtol=1e-11, atol=1e-9), (i, key, a, b)
    assert n_lvl > 50 and res["in_image"].sum() > 500


@pytest.mark.parametrize("estimator", [0, 1, 2])
def test_pose_update_parity(capi, ora, estimator):
    """Tracker pose update: CalcJacobian per point (1e-11) and CalcPoseUpdate (robust WLS<6>, 1e-9) vs the oracle."""
    rng = np.random.default_rng(17 + estimator)
    cams, extr = synth.make_rig(2, rng)
    cam = cams[1]
    f = capi.FeHandle(640, 480)
    f.set_camera(cam)
    B = synth.rt_pack((synth.so3_exp(rng.normal(0, 0.2, 3)), rng.normal(0, 0.5, 3)))
    Cb = synth.rt_pack(extr[1])
    n = 1500
    Rb, tb = B[:9].reshape(3, 3), B[9:]
    Rc, tc = Cb[:9].reshape(3, 3), Cb[9:]
    pc = rng.normal(0, 1, (n, 3)) * [3, 3, 1.5] + [0, 0, 6]
    pw = ((pc - tc) @ Rc - tb) @ Rb                      # world = Rb^T (Rc^T (pc - tc) - tb)
    jr = f.calc_jacobians(B, Cb, pw)
    meas = np.zeros(n, capi.POSE_MEAS_DTYPE)
    for i in range(n):
        px, D, J, inv = ora.calc_jacobian(cam, B, Cb, pw[i])
        assert np.allclose(jr[i]["px"], px, rtol=1e-12, atol=1e-9)
        assert np.allclose(jr[i]["jac"].reshape(2, 6), J, rtol=1e-10, atol=1e-8), i
    meas["image"] = jr["px"]; meas["jac"] = jr["jac"]
    lvl = rng.integers(0, 4, n)
    meas["sqrt_inv_noise"] = 1.0 / (1 << lvl)
    true_mu = np.array([0.02, -0.01, 0.015, 0.004, -0.003, 0.002])
    pred = np.einsum("nij,j->ni", jr["jac"].reshape(n, 2, 6), true_mu)
    meas["found"] = jr["px"] + pred + rng.normal(0, 0.3, (n, 2)) * (1 << lvl)[:, None]
    bad = rng.random(n) < 0.1
    meas["found"][bad] += rng.uniform(-40, 40, (bad.sum(), 2))
    meas["found_flag"] = (rng.random(n) < 0.8) & (jr["in_image"] == 1)
    for override in (0.0, 25.0):
        res, outl = f.pose_update(meas, estimator, override)
        mu, sig, out_ref, nin = ora.pose_update(meas["found"], meas["image"], meas["sqrt_inv_noise"], meas["jac"], meas["found_flag"], estimator, override)
        assert abs(res["sigma_sq"] - sig) <= 1e-12 * sig         # same median element (errors differ by FMA rounding only)
        assert np.array_equal(outl, out_ref) and res["n_inliers"] == nin
        assert np.linalg.norm(res["mu"] - mu) <= 1e-9 * np.linalg.norm(mu)
    assert np.linalg.norm(res["mu"] - true_mu) < 0.01       # and it recovers the motion
    empty = np.zeros(5, capi.POSE_MEAS_DTYPE)
    res, _ = f.pose_update(empty, estimator)
    assert not res["mu"].any() and res["n_valid"] == 0      # no valid measurements -> null update (src/Tracker.cc:1423-1424)


def test_fe_errors(capi):
    f = capi.FeHandle(640, 480)
    req = np.zeros(1, capi.PATCH_REQ_DTYPE)
    with pytest.raises(capi.McpError):
        f.search_patches(0, req)                                  # empty keyframe slot
    with pytest.raises(capi.McpError):
        f.make_keyframe(99, np.zeros((480, 640), np.uint8))


@pytest.mark.parametrize("mode", ["fast_percent", "shi_percent", "fast_thresh", "strict", "pruned", "pruned_shi_thresh"])
def test_keyframe_rest_candidates(capi, ora, mode):
    """MakeKeyFrame_Rest candidate generation (src/KeyFrame.cc:363-531): fast_nonmax, scoring, selection order and
    the MiniPatch stable-point test — candidates index-exact and scores bit-exact against the oracle, all 4 levels."""
    cur = synth.make_frame(seed=11)
    prev = synth.make_frame(seed=11, shift=(1.0, -1.0))
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    f.make_keyframe(0, prev)
    f.make_keyframe(1, cur)
    kw = dict(use_shi=mode in ("shi_percent", "pruned_shi_thresh"), use_thresh=mode in ("fast_thresh", "pruned_shi_thresh"),
              nonmax_strict=mode == "strict", thresh=40.0 if mode == "fast_thresh" else 500.0)
    pruned = mode.startswith("pruned")
    got = f.make_keyframe_rest(1, prev_slot=0 if pruned else -1, n_prev=2 if pruned else 0,
                               **{k: int(v) if isinstance(v, bool) else v for k, v in kw.items()})
    pc, pp = ora.pyramid(cur), ora.pyramid(prev)
    total = 0
    for l in range(4):
        lc, lp = ora.level_corners(pc[l]), ora.level_corners(pp[l])
        ref = ora.keyframe_rest_level(pc[l], lc, prev_img=pp[l] if pruned else None, prev_lev=lp if pruned else None,
                                      n_prev=2 if pruned else 0, **kw)
        g = got[l]
        assert g["n_max"] == ref["n_max"], (l, g["n_max"], ref["n_max"])
        assert g["n_candidates"] == ref["n_candidates"], (l, g["n_candidates"], ref["n_candidates"])
        assert np.array_equal(np.stack([g["cand"]["x"], g["cand"]["y"]], 1), ref["xy"])
        assert np.array_equal(g["cand"]["score"], ref["score"])
        total += ref["n_candidates"]
    assert total > 50


def test_glare_masking(capi, ora):
    """bGlareMasking (src/KeyFrame.cc:214-242): the per-level lastMask equals the oracle's (cv2-pinned) glare mask AND the
    internal mask, and the corners are the oracle's corners filtered with that mask."""
    img = synth.make_frame(seed=41).copy()
    rng = np.random.default_rng(41)
    for _ in range(25):
        x, y = int(rng.integers(0, 640)), int(rng.integers(0, 480))
        img[max(y - 3, 0):y + 4, max(x - 5, 0):x + 6] = 255
    internal = np.full((480, 640), 255, np.uint8)
    internal[:, :64] = 0
    f = capi.FeHandle(640, 480, max_corners_per_level=16384)
    for use_internal in (False, True):
        f.set_mask(internal if use_internal else None)
        f.set_glare_masking(True)
        lv = f.make_keyframe(0, img, want_images=True, want_masks=True)
        pyr = ora.pyramid(img)
        mpyr = ora.pyramid(internal) if use_internal else [None] * 4
        for l in range(4):
            want_mask = ora.glare_mask(pyr[l], mpyr[l])
            assert np.array_equal(lv[l]["last_mask"], want_mask), l
            ref = ora.level_corners(pyr[l], mask=want_mask)
            assert lv[l]["fast_thresh"] == ref["fast_thresh"] and np.array_equal(lv[l]["corners"], ref["corners"]), l
            assert l > 0 or (want_mask == 0).sum() > 0            # (the blobs average out below 245 on the coarse levels)
        f.set_glare_masking(False)
        lv0 = f.make_keyframe(0, img, want_masks=True)
        assert np.array_equal(lv0[0]["last_mask"], internal if use_internal else np.full((480, 640), 255, np.uint8))
        assert lv0[0]["n_corners"] >= lv[0]["n_corners"]
