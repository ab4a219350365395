"""Pins the CPU oracle (oracle/*.c, a restatement) against the REFERENCE'S OWN code: oracle/_ref/libref.so is
include/mcptam/MEstimator.h, src/ShiTomasi.cc, src/MiniPatch.cc, src/TaylorCamera.cc and src/PatchFinder.cc (SSE and
scalar ZMSSD branches) compiled from /root/reference by oracle/build_ref.py against minimal TooN/libCVD/ROS stand-ins.

Integer results (scores, templates, corner picks) must be bit-identical; fp64 results of the same formulae are compared
bit-for-bit where the operation order is the reference's, otherwise to a few ulp (stated per test).
SURVEY.md §8 rows pinned here: a1-a4, a9, a15 (sigma), a20, a21-a28.
"""
import ctypes as C

import numpy as np
import pytest

from mcptam_b200 import synth
from oracle import oracle as ora
from oracle import ref as R

pytestmark = pytest.mark.skipif(R.lib() is None, reason="neither /root/reference nor a prebuilt oracle/_ref/libref.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [7, 8, 101, 4096, 80157])
def test_mestimator_sigmas(n):
    """Huber / Tukey sigma^2 (upper median [n/2], 1 + 5 / (2n - 6) in size_t arithmetic): MEstimator.h:109-126,194-204."""
    rng = np.random.default_rng(n)
    e2 = np.ascontiguousarray(rng.gamma(1.5, 3.0, n) * rng.choice([1.0, 40.0], n, p=[0.95, 0.05]))
    L, O = R.lib(), ora.lib()
    assert L.ref_huber_sigma_sq(_p(e2), n) == O.ora_huber_sigma_sq(_p(e2), n)
    assert L.ref_tukey_sigma_sq(_p(e2), n) == O.ora_tukey_sigma_sq(_p(e2), n)


def test_level_helpers():
    L = R.lib()
    for lvl in range(4):
        for p in (0.0, 3.0, 17.25, 639.0):
            assert L.ref_level_zero_pos(p, lvl) == (p + 0.5) * (1 << lvl) - 0.5
            assert L.ref_level_n_pos(p, lvl) == (p + 0.5) / (1 << lvl) - 0.5


def test_shitomasi():
    img = synth.make_frame(seed=3)
    rng = np.random.default_rng(0)
    L = R.lib()
    for _ in range(300):
        x, y = int(rng.integers(5, 635)), int(rng.integers(5, 475))
        assert L.ref_shitomasi(_p(img), 640, 480, 640, 3, x, y) == ora.shitomasi(img, x, y)


def test_minipatch():
    a = synth.make_frame(seed=11)
    b = synth.make_frame(seed=11, shift=(2.0, -1.0))
    lv = ora.level_corners(b)
    cor, lut = np.ascontiguousarray(lv["corners"], np.int32), np.ascontiguousarray(lv["row_lut"], np.int32)
    src = ora.level_corners(a)["corners"]
    src = src[(src[:, 0] > 12) & (src[:, 0] < 628) & (src[:, 1] > 12) & (src[:, 1] < 468)]
    rng = np.random.default_rng(1)
    L = R.lib()
    n_found = 0
    for k in rng.choice(len(src), 200, replace=False):
        sx, sy = int(src[k, 0]), int(src[k, 1])
        for use_lut in (True, False):
            pos = np.array([sx - 2, sy + 1], np.int32)
            f_ref = L.ref_minipatch_find(_p(a), _p(b), 640, 480, 640, sx, sy, _p(pos), 10, _p(cor), len(cor), _p(lut) if use_lut else None)
            f_ora, pos_o = ora.minipatch_find(a, b, cor, lut, (sx, sy), (sx - 2, sy + 1), 10)
            assert bool(f_ref) == f_ora
            if f_ora:
                assert tuple(pos) == tuple(pos_o)
                n_found += 1
    assert n_found > 100


def _cams():
    rng = np.random.default_rng(5)
    out = []
    for _ in range(3):
        a0, a2, a3, a4 = synth.DEFAULT_TAYLOR
        params = [a0 * (1 + 0.01 * rng.standard_normal()), a2, a3, a4, 320 + 2 * rng.standard_normal(), 240 + 2 * rng.standard_normal(),
                  1.0 + 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal()]
        out.append(params)
    return out


@pytest.mark.parametrize("k", [0, 1, 2])
def test_taylor_camera_refresh_params(k):
    """RefreshParams / FindInvPolyUsingRoots (src/TaylorCamera.cc:84-198, 489-604) vs the numpy restatement that fills the
    C-ABI camera struct: exact for the closed-form members; the inverse polynomial is a least-squares fit through a different
    SVD, so its VALUES (rho over the valid theta range) are compared, to 1e-7 px."""
    params = _cams()[k]
    cam = synth.taylor_camera(params)
    d = R.RefCamera(params).derived()
    assert np.array_equal(d["center"], np.array(cam.center)) and np.array_equal(d["affine"], np.array(cam.affine))
    assert d["min_theta"] == cam.min_theta
    assert abs(d["theta_mean"] - cam.theta_mean) <= 1e-12 and abs(d["theta_std"] - cam.theta_std) <= 1e-12
    assert len(d["inv_poly"]) == cam.n_inv
    th = np.linspace(cam.min_theta, np.pi / 2 - 0.002, 2000)
    xs_r = (th - d["theta_mean"]) / d["theta_std"]; xs_o = (th - cam.theta_mean) / cam.theta_std
    rho_r = np.polynomial.polynomial.polyval(xs_r, d["inv_poly"]); rho_o = np.polynomial.polynomial.polyval(xs_o, np.array(cam.inv_poly[:cam.n_inv]))
    assert np.abs(rho_r - rho_o).max() < 1e-7


@pytest.mark.parametrize("k", [0, 1])
def test_taylor_camera_project_derivs_unproject(k):
    """Project / GetProjectionDerivs / GetCamSphereDeriv / UnProject (src/TaylorCamera.cc:202-383, 617-669) vs the oracle's C
    restatement on the SAME derived parameters (taken from the reference object): bit-identical."""
    params = _cams()[k]
    rc = R.RefCamera(params)
    d = rc.derived()
    cam = synth.taylor_camera(params)
    cam.theta_mean, cam.theta_std, cam.min_theta = d["theta_mean"], d["theta_std"], d["min_theta"]
    for i, v in enumerate(d["inv_poly"]):
        cam.inv_poly[i] = float(v)
    rng = np.random.default_rng(k)
    pts = rng.normal(0, 1, (4000, 3)) * [4, 4, 3] + [0, 0, 4]
    pts[:5] = [[0, 0, 3], [0, 0, -2], [1e-9, 0, 1], [0, 2, 0], [-3, 1e-12, 0.5]]        # pole, behind, on the axes
    px, inv, D, dt, dp = rc.project(pts)
    O = ora.lib()
    O.ora_cam_project.argtypes = [C.c_void_p] * 4
    O.ora_cam_sphere_deriv.argtypes = [C.c_void_p] * 3
    O.ora_cam_unproject.argtypes = [C.c_void_p] * 3
    for i in range(len(pts)):
        p = np.ascontiguousarray(pts[i]); px_o = np.zeros(2); D_o = np.zeros(4); dt_o = np.zeros(3); dp_o = np.zeros(3)
        inv_o = O.ora_cam_project(C.byref(cam), _p(p), _p(px_o), _p(D_o))
        O.ora_cam_sphere_deriv(_p(p), _p(dt_o), _p(dp_o))
        assert inv_o == inv[i], (i, pts[i])
        assert np.array_equal(px_o, px[i]) and np.array_equal(dt_o, dt[i]) and np.array_equal(dp_o, dp[i]), (i, pts[i], px_o, px[i])
        assert np.array_equal(D_o, D[i], equal_nan=True), (i, D_o, D[i])
    pix = rng.uniform([0, 0], [640, 480], (2000, 2))
    rays = rc.unproject(pix)
    for i in range(len(pix)):
        q = np.ascontiguousarray(pix[i]); r_o = np.zeros(3)
        O.ora_cam_unproject(C.byref(cam), _p(q), _p(r_o))
        assert np.array_equal(r_o, rays[i])


def test_zmssd_sse_and_scalar():
    """PatchFinder::ZMSSDAtPoint, both branches of the reference (src/PatchFinder.cc:511-658), vs the oracle."""
    img = synth.make_frame(seed=21)
    rng = np.random.default_rng(2)
    O = ora.lib()
    for _ in range(400):
        t = np.ascontiguousarray(rng.integers(0, 256, 64).astype(np.uint8)) if rng.random() < 0.3 else None
        x, y = int(rng.integers(0, 640)), int(rng.integers(0, 480))
        if t is None:
            tx, ty = int(rng.integers(8, 630)), int(rng.integers(8, 470))
            t = np.ascontiguousarray(img[ty - 4:ty + 4, tx - 4:tx + 4]).reshape(-1)
        ts, tq = int(t.astype(np.int64).sum()), int((t.astype(np.int64) ** 2).sum())
        want = O.ora_zmssd(_p(img), 640, 480, 640, _p(t), ts, tq, x, y, 64 * 250)
        assert R.lib().ref_zmssd(_p(img), 640, 480, 640, _p(t), x, y) == want
        assert R.lib(scalar=True).ref_zmssd(_p(img), 640, 480, 640, _p(t), x, y) == want


def test_calc_search_level_and_warp_matrix():
    """PatchFinder::CalcSearchLevelAndWarpMatrix (src/PatchFinder.cc:69-122) vs ora_project_point's warp / level."""
    rng = np.random.default_rng(9)
    cams, extr = synth.make_rig(2, rng)
    cam = cams[0]
    L = R.lib()
    n_ok = 0
    for _ in range(500):
        T = synth.rt_pack((synth.so3_exp(rng.normal(0, 0.3, 3)), rng.normal(0, 0.5, 3)))
        pw = rng.normal(0, 1, 3) * [3, 3, 2] + [0, 0, 6]
        right = rng.normal(0, 1, 3) * 0.01 * rng.choice([0.3, 1.0, 3.0]); down = np.cross([0, 0, 1.0], right) + rng.normal(0, 1e-3, 3)
        o = ora.project_point(cam, T, pw, right, down)
        W = np.zeros(4)
        lvl = L.ref_calc_search_level(_p(T), _p(np.ascontiguousarray(pw)), _p(np.ascontiguousarray(right)), _p(np.ascontiguousarray(down)),
                                      _p(np.ascontiguousarray(o["derivs"])), _p(W))
        assert lvl == o["level"]
        assert np.array_equal(W, o["warp_inv"]), (W, o["warp_inv"])
        n_ok += lvl >= 0
    assert n_ok > 50


@pytest.mark.parametrize("scalar", [False, True])
def test_patch_search_sequence(scalar):
    """MakeTemplateCoarseCont -> FindPatchCoarse -> MakeSubPixTemplate -> IterateSubPixToConvergence on the reference's
    PatchFinder vs the oracle: template bytes, found flags, coarse position, score and the sub-pixel position (bit-exact: both
    are compiled without FMA contraction)."""
    a = synth.make_frame(seed=31)
    b = synth.make_frame(seed=31, shift=(3.0, -2.0))
    pyr_a, pyr_b = ora.pyramid(a), ora.pyramid(b)
    lv_b = [ora.level_corners(im) for im in pyr_b]
    cor = ora.level_corners(a)["corners"]
    cor = cor[(cor[:, 0] > 20) & (cor[:, 0] < 620) & (cor[:, 1] > 20) & (cor[:, 1] < 460)]
    rng = np.random.default_rng(4)
    n_found = n_sub = n_bad = 0
    for k in rng.choice(len(cor), 400, replace=False):
        lvl = int(rng.choice([0, 0, 1, 2]))
        s = float(1 << lvl)
        ang = rng.normal(0, 0.05)
        req = dict(src_level=0, src_cx=int(cor[k, 0]), src_cy=int(cor[k, 1]), search_level=lvl,
                   warp_inv=np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]).reshape(-1) * s * rng.uniform(0.9, 1.1),
                   pred_x=int(cor[k, 0] - 3 + rng.integers(-2, 3)), pred_y=int(cor[k, 1] + 2 + rng.integers(-2, 3)),
                   range=int(rng.choice([6, 10, 15])), subpix_its=int(rng.choice([0, 8, 10])), exhaustive=int(rng.random() < 0.1))
        if rng.random() < 0.05:
            req["src_cx"], req["src_cy"] = 3, 2                                           # template leaves the source image
        want = ora.search_patch(pyr_a, pyr_b, lv_b, req)
        got = R.patch_search(pyr_a[0], req, pyr_b[lvl], lv_b[lvl]["corners"], lv_b[lvl]["row_lut"], scalar=scalar)
        assert np.array_equal(got["template"], want["template"])
        assert got["template_bad"] == want["template_bad"]
        n_bad += got["template_bad"]
        if got["template_bad"]:
            continue
        assert got["score"] == want["score"] and got["found"] == want["found"], (req, got, want)
        if want["found"]:                                                # GetCoarsePos(): LevelZeroPos of the best corner, truncated
            ls = 1 << lvl
            assert (got["coarse_x"], got["coarse_y"]) == (int((want["coarse_x"] + 0.5) * ls - 0.5), int((want["coarse_y"] + 0.5) * ls - 0.5))
        if want["found"]:
            assert got["found_x"] == want["found_x"] and got["found_y"] == want["found_y"], (req, got, want)
            n_found += 1
            n_sub += got["did_subpix"]
    assert n_found > 150 and n_sub > 80 and n_bad > 5


# ---------------------------------------------------------------------------------------------------------------------
# Bundle adjuster: the reference's src/ChainBundle.cc (compiled unmodified, g2o replaced by a dense stand-in) vs the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _ref_problem(name_or_prob, seed=0, variant=""):
    """The synthetic problem with its camera structs re-derived by the REFERENCE's TaylorCamera (inverse polynomial, theta
    mean / std), so that oracle and reference evaluate exactly the same camera model; plus the 9 parameters per camera."""
    import copy
    prob = synth.make_ba_config(name_or_prob, seed=seed) if isinstance(name_or_prob, str) else name_or_prob
    if "fixed" in variant:
        prob = synth.with_fixed_points(prob, 0.15)
    if "single" in variant:
        prob = synth.as_single_link(prob)
    prob = copy.copy(prob)
    params, cams = [], []
    for cam in prob.cams:
        p9 = [cam.poly[0], cam.poly[2], cam.poly[3], cam.poly[4], cam.center[0], cam.center[1], cam.affine[0], cam.affine[1], cam.affine[2]]
        d = R.RefCamera(p9).derived()
        c2 = synth.TaylorCamStruct.from_buffer_copy(bytes(cam))
        assert np.array_equal(d["center"], np.array(c2.center)) and np.array_equal(d["affine"], np.array(c2.affine)) and len(d["inv_poly"]) == c2.n_inv
        c2.theta_mean, c2.theta_std, c2.min_theta = d["theta_mean"], d["theta_std"], d["min_theta"]
        for i, v in enumerate(d["inv_poly"]):
            c2.inv_poly[i] = float(v)
        params.append(p9); cams.append(c2)
    prob.cams = cams
    return prob, params


@pytest.mark.parametrize("cfg,seed,variant", [("tiny", 0, ""), ("tiny", 3, "fixed"), ("tiny", 1, "single"), ("tiny", 2, "fixed+single"), ("cfg1", 0, ""), ("cfg1", 2, "fixed")])
def test_chainbundle_residuals_jacobians_oplus(cfg, seed, variant):
    """EdgeChainMeas::computeError / chi2 / linearizeOplus, PoseChainHelper::UpdateTransforms, Vertex*::oplusImpl and the adaptive
    Huber kernel of the reference vs the oracle: same formulae in the same order -> compared to 1e-12 relative (residuals,
    chi2, sigma^2, rho) and 1e-9 (Jacobians: the reference recomputes intermediate transforms per column)."""
    from oracle.oracle import OracleBA
    prob, params = _ref_problem(cfg, seed, variant)
    r = R.RefBA(prob, params)
    o = OracleBA(prob)
    er, cr = r.eval()
    eo, co = o.eval()
    assert np.allclose(er, eo, rtol=1e-12, atol=1e-10) and np.allclose(cr, co, rtol=1e-11, atol=1e-12)
    if "fixed" in variant:
        assert (cr < 0).sum() > 0 and np.array_equal(cr < 0, co < 0)
    rng = np.random.default_rng(seed)
    for m in rng.choice(prob.n_meas, min(prob.n_meas, 150), replace=False):
        jo_r, js_r, jp_r = r.jacobians(m)
        jo_o, js_o, jp_o = o.jacobians(m)
        for a, b in ((jo_r, jo_o), (js_r, js_o), (jp_r, jp_o)):
            assert np.abs(np.asarray(a) - np.asarray(b)).max() <= 1e-9 * max(np.abs(b).max(), 1.0), (m, a, b)
    P0, X0 = o.poses().copy(), o.points().copy()
    for i in np.flatnonzero(prob.pose_fixed == 0)[:4]:
        d = rng.normal(0, 0.05, 6)
        o.oplus_pose(i, d)                                             # the oracle applies in place
        assert np.allclose(r.oplus_pose(i, d), o.poses()[i], rtol=1e-13, atol=1e-14)
    for p in np.flatnonzero(prob.pt_fixed == 0)[:20]:
        d = rng.normal(0, 0.01, 3)
        o.oplus_point(p, d)
        assert np.allclose(r.oplus_point(p, d), o.points()[p], rtol=1e-12, atol=1e-13)
    o.set_state(P0, X0)
    eo, co = o.eval()
    rho, sig = r.robustify()
    L = ora.lib()
    a = np.ascontiguousarray(np.abs(co))
    assert sig == L.ora_huber_sigma_sq(_p(a), len(a))


@pytest.mark.parametrize("seed,variant,robust", [(0, "", True), (2, "", True), (3, "fixed", True), (1, "single", True), (0, "", False)])
def test_chainbundle_compute(seed, variant, robust):
    """ChainBundle::Compute of the reference (its actions, robust kernel, convergence tests, Tukey outlier pass, return codes;
    LM loop and linear solve from the dense g2o stand-in) vs the oracle: same iteration / trial counts, lambda, sigma^2,
    outlier set, final poses and points to 1e-7 relative."""
    from oracle.oracle import OracleBA
    prob, params = _ref_problem("tiny", seed, variant)
    r = R.RefBA(prob, params, use_robust=robust, use_tukey=robust)
    o = OracleBA(prob, use_robust=robust, use_tukey=robust)
    for n_iter in (6, 5):                                          # Compute twice on the same object (two-step adjuster)
        rc_r, st_r = r.compute(n_iter)
        rc_o, st_o = o.compute(n_iter)
        assert rc_r == rc_o
        assert st_r["total_trials"] == st_o.total_trials and st_r["converged"] == st_o.converged
        if robust:                       # (without a robust kernel the reference never writes RobustKernelData::_dSigmaSquared: uninitialised)
            assert abs(st_r["sigma_sq"] - st_o.sigma_sq) <= 1e-7 * st_o.sigma_sq
        assert abs(st_r["lambda_"] - st_o.lambda_) <= 1e-6 * st_o.lambda_
        assert abs(st_r["mean_chi2"] - st_o.mean_chi2) <= 1e-7 * abs(st_o.mean_chi2)
        P, X = r.state()
        assert np.linalg.norm(P - o.poses()) <= 1e-7 * np.linalg.norm(o.poses())
        assert np.linalg.norm(X - o.points()) <= 1e-7 * np.linalg.norm(o.points())
        assert r.outliers().tolist() == sorted(o.outliers().tolist())


def test_chainbundle_marginals_and_return_codes():
    """< 3 movable poses: the median point-depth covariance (GetMaxCov, src/ChainBundle.cc:1401-1448)."""
    from oracle.oracle import OracleBA
    base = synth.make_ba_problem(n_cam=2, n_mkf=3, n_pt=60, seed=3)
    prob, params = _ref_problem(base)
    r = R.RefBA(prob, params)
    o = OracleBA(prob)
    rc_r, st_r = r.compute(5)
    rc_o, st_o = o.compute(5)
    assert rc_r == rc_o and st_r["total_trials"] == st_o.total_trials
    assert st_o.max_cov > 0 and abs(st_r["max_cov"] - st_o.max_cov) <= 1e-6 * st_o.max_cov
