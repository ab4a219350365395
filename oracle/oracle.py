"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Pinned against the reference's own code where it exists (oracle/_ref, tests/test_oracle_vs_ref.py); third-party behaviour restated (see oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCP_ORACLE_FAST=1 (set by bench.py for its CPU timing legs only): the -O3 -march=native build.  It is compiled on the
# machine that runs it and tagged with that machine's CPU flags, because a prebuilt file may travel to another host.
_FAST = os.environ.get("MCP_ORACLE_FAST", "0") == "1"


def _cpu_tag() -> str:
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_LIB_PATH = os.path.join(_HERE, "_build", ("liboracle_fast_%s.so" % _cpu_tag()) if _FAST else "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("ba_oracle.c", "fe_oracle.c", "oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        cmd = ["make", "-C", _HERE, "-B"]
        if _FAST:
            cmd += ["FAST=1", "OUT=" + os.path.relpath(_LIB_PATH, _HERE)]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return _LIB_PATH


class BaStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("total_trials", C.c_int32), ("converged", C.c_int32),
                ("hit_max_iter", C.c_int32), ("n_outliers", C.c_int32), ("pad_", C.c_int32),
                ("sigma_sq", C.c_double), ("mean_chi2", C.c_double), ("lambda_", C.c_double),
                ("max_cov", C.c_double), ("chi2_before", C.c_double), ("chi2_after", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.ora_ba_create.restype = C.c_void_p
        L.ora_ba_create.argtypes = [C.c_int, C.c_int]
        L.ora_ba_destroy.argtypes = [C.c_void_p]
        L.ora_ba_set_cameras.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_ba_load.argtypes = [C.c_void_p] + [C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5
        L.ora_ba_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.POINTER(BaStats)]
        for f in ("ora_ba_get_poses", "ora_ba_get_points", "ora_ba_set_poses", "ora_ba_set_points"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        L.ora_ba_get_outliers.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ora_ba_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_ba_jacobians.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_ba_oplus_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_ba_oplus_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_ba_lm_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_huber_sigma_sq.restype = C.c_double
        L.ora_huber_sigma_sq.argtypes = [C.c_void_p, C.c_int]
        L.ora_tukey_sigma_sq.restype = C.c_double
        L.ora_tukey_sigma_sq.argtypes = [C.c_void_p, C.c_int]
        L.ora_cam_project.argtypes = [C.c_void_p] * 4
        L.ora_cam_sphere_deriv.argtypes = [C.c_void_p] * 3
        L.ora_cam_unproject.argtypes = [C.c_void_p] * 3
        L.ora_se3_exp.argtypes = [C.c_void_p] * 2
        L.ora_so3_exp.argtypes = [C.c_void_p] * 2
        # front end
        L.ora_halfsample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        for f in ("ora_fast10_detect", "ora_fast10_detect_bruteforce"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ora_fastN_detect_bruteforce.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        for f in ("ora_fast10_score", "ora_fast10_score_bisect"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_level_corners.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_shitomasi.restype = C.c_double
        L.ora_shitomasi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ora_patch_template.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.ora_zmssd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 5
        L.ora_find_patch_coarse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
        L.ora_subpix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ora_minipatch_ssd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.ora_minipatch_find.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ora_project_point.argtypes = [C.c_void_p] * 10
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def cam_array(cams):
    arr = (type(cams[0]) * len(cams))()
    for i, c in enumerate(cams):
        C.memmove(C.byref(arr[i]), C.byref(c), C.sizeof(c))
    return arr


class OracleBA:
    """Thin object wrapper over the ora_ba_* functions."""

    def __init__(self, prob, use_robust=True, use_tukey=True):
        self.L = lib()
        self.h = C.c_void_p(self.L.ora_ba_create(int(use_robust), int(use_tukey)))
        self.prob = prob
        self._cams = cam_array(prob.cams)
        self.L.ora_ba_set_cameras(self.h, len(prob.cams), C.cast(self._cams, C.c_void_p))
        self._keep = [np.ascontiguousarray(prob.pose_Rt, np.float64), np.ascontiguousarray(prob.pose_fixed, np.uint8),
                      np.ascontiguousarray(prob.pt_xyz, np.float64), np.ascontiguousarray(prob.pt_chain, np.int32),
                      np.ascontiguousarray(prob.pt_fixed, np.uint8), np.ascontiguousarray(prob.meas_xy, np.float64),
                      np.ascontiguousarray(prob.meas_chain, np.int32), np.ascontiguousarray(prob.meas_pt, np.int32),
                      np.ascontiguousarray(prob.meas_noise, np.float64), np.ascontiguousarray(prob.meas_cam, np.int32)]
        k = self._keep
        rc = self.L.ora_ba_load(self.h, prob.n_pose, _p(k[0]), _p(k[1]), prob.n_pt, _p(k[2]), _p(k[3]), _p(k[4]),
                                prob.n_meas, _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]))
        if rc != 0:
            raise RuntimeError("ora_ba_load failed: %d" % rc)
        self.n_pose_var = int((np.asarray(prob.pose_fixed) == 0).sum())
        self.n_pt_var = int((np.asarray(prob.pt_fixed) == 0).sum())

    def __del__(self):
        try:
            self.L.ora_ba_destroy(self.h)
        except Exception:
            pass

    def compute(self, n_iter=100, user_lambda=-1.0, solve_mode=0, abort=None):
        st = BaStats()
        ab = np.zeros(1, np.uint8) if abort is None else abort
        rc = self.L.ora_ba_compute(self.h, _p(ab), n_iter, float(user_lambda), solve_mode, C.byref(st))
        return rc, st

    def poses(self):
        o = np.zeros((self.prob.n_pose, 12))
        self.L.ora_ba_get_poses(self.h, _p(o))
        return o

    def points(self):
        o = np.zeros((self.prob.n_pt, 3))
        self.L.ora_ba_get_points(self.h, _p(o))
        return o

    def set_state(self, poses, points):
        poses = np.ascontiguousarray(poses, np.float64)
        points = np.ascontiguousarray(points, np.float64)
        self.L.ora_ba_set_poses(self.h, _p(poses))
        self.L.ora_ba_set_points(self.h, _p(points))

    def outliers(self):
        o = np.zeros(self.prob.n_meas, np.int32)
        n = self.L.ora_ba_get_outliers(self.h, _p(o), len(o))
        return o[:n].copy()

    def eval(self):
        e = np.zeros((self.prob.n_meas, 2))
        c = np.zeros(self.prob.n_meas)
        self.L.ora_ba_eval(self.h, _p(e), _p(c))
        return e, c

    def jacobians(self, m):
        jo = np.zeros((2, 2, 6)); js = np.zeros((2, 2, 6)); jp = np.zeros((2, 3))
        self.L.ora_ba_jacobians(self.h, int(m), _p(jo), _p(js), _p(jp))
        return jo, js, jp

    def oplus_pose(self, i, d):
        d = np.ascontiguousarray(d, np.float64)
        self.L.ora_ba_oplus_pose(self.h, int(i), _p(d))

    def oplus_point(self, i, d):
        d = np.ascontiguousarray(d, np.float64)
        self.L.ora_ba_oplus_point(self.h, int(i), _p(d))

    def lm_step(self, lam, sigma_sq=-1.0, solve_mode=0):
        d = np.zeros(6 * self.n_pose_var + 3 * self.n_pt_var)
        s = C.c_double(); r = C.c_double()
        rc = self.L.ora_ba_lm_step(self.h, float(lam), float(sigma_sq), solve_mode, _p(d), C.byref(s), C.byref(r))
        return rc, d, s.value, r.value


# ---------------------------------------------------------------------------------------------
# front-end wrappers (numpy in / numpy out)
# ---------------------------------------------------------------------------------------------
def halfsample(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.uint8)
    lib().ora_halfsample(_p(img), w, h, w, _p(out), w // 2)
    return out


def pyramid(img, levels=4):
    out = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(levels - 1):
        out.append(halfsample(out[-1]))
    return out


def fast10_detect(img, b, brute=False, n_arc=10):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    L = lib()
    if brute:
        n = L.ora_fastN_detect_bruteforce(_p(img), w, h, w, b, n_arc, None, 0)
        xy = np.zeros((max(n, 1), 2), np.int32)
        L.ora_fastN_detect_bruteforce(_p(img), w, h, w, b, n_arc, _p(xy), n)
    else:
        n = L.ora_fast10_detect(_p(img), w, h, w, b, None, 0)
        xy = np.zeros((max(n, 1), 2), np.int32)
        L.ora_fast10_detect(_p(img), w, h, w, b, _p(xy), n)
    return xy[:n]


def fast10_score(img, xy, b, bisect=False):
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32)
    sc = np.zeros(max(len(xy), 1), np.int32)
    f = lib().ora_fast10_score_bisect if bisect else lib().ora_fast10_score
    f(_p(img), img.shape[1], _p(xy), len(xy), b, _p(sc))
    return sc[:len(xy)]


def level_corners(img, mask=None, adaptive=True, fixed_thresh=10, cap=1 << 20):
    """One pyramid level of MakeKeyFrame_Lite."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    xy = np.zeros((cap, 2), np.int32)
    freq = np.zeros(31, np.int32)
    thr = C.c_int32()
    lut = np.zeros(h, np.int32)
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    n = lib().ora_level_corners(_p(img), w, h, w, _p(m), w if m is not None else 0, int(adaptive), fixed_thresh, _p(xy), cap,
                                _p(freq), C.byref(thr), _p(lut))
    return {"corners": xy[:n].copy(), "n_corners": n, "fast_freq": freq, "fast_thresh": thr.value, "row_lut": lut}


def fast_nonmax(img, corners, barrier, strict=False):
    """[3P] CVD::fast_nonmax restatement: boolean keep mask over `corners`."""
    img = np.ascontiguousarray(img, np.uint8)
    corners = np.ascontiguousarray(corners, np.int32)
    keep = np.zeros(len(corners) + 1, np.uint8)
    h, w = img.shape
    f = lib().ora_fast_nonmax
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    f(_p(img), w, h, w, _p(corners), len(corners), int(barrier), int(strict), _p(keep))
    return keep[:len(corners)].astype(bool)


def keyframe_rest_level(img, lev, prev_img=None, prev_lev=None, n_prev=0, use_shi=False, use_thresh=False, top_fraction=0.8,
                        thresh=70.0, nonmax_strict=False):
    """One level of KeyFrame::MakeKeyFrame_Rest (src/KeyFrame.cc:363-531).  lev / prev_lev: dicts from level_corners()."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cor = np.ascontiguousarray(lev["corners"], np.int32)
    lut = np.ascontiguousarray(lev["row_lut"], np.int32)
    cap = len(cor) + 1
    out_xy = np.zeros((cap, 2), np.int32)
    out_sc = np.zeros(cap)
    n_max = C.c_int32()
    f = lib().ora_keyframe_rest_level
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                  C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    if prev_img is not None and n_prev > 0:
        pim = np.ascontiguousarray(prev_img, np.uint8)
        pcor = np.ascontiguousarray(prev_lev["corners"], np.int32)
        plut = np.ascontiguousarray(prev_lev["row_lut"], np.int32)
        pa = (_p(pim), _p(pcor), len(pcor), _p(plut), int(n_prev))
    else:
        pim = pcor = plut = None
        pa = (None, None, 0, None, 0)
    n = f(_p(img), w, h, w, _p(cor), len(cor), _p(lut), int(lev["fast_thresh"]), int(use_shi), int(use_thresh), float(top_fraction),
          float(thresh), int(nonmax_strict), *pa, _p(out_xy), _p(out_sc), cap, C.byref(n_max))
    return {"n_max": n_max.value, "n_candidates": n, "xy": out_xy[:n].copy(), "score": out_sc[:n].copy()}


def glare_mask(img, internal=None):
    """Level::lastMask with bGlareMasking (src/KeyFrame.cc:214-242): five dilations by OpenCV's 5x5 MORPH_ELLIPSE element
    (rows +-2: the centre pixel, rows -1..1: five pixels; out-of-image pixels never count), threshold 245 inverted, AND with the
    internal mask.  [3P] OpenCV semantics restated by definition (iterated max filter); pinned against cv2 in tests/test_oracle_cpu.py."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    cur = img.copy()
    offs = [(0, -2), (0, 2)] + [(dx, dy) for dy in (-1, 0, 1) for dx in (-2, -1, 0, 1, 2)]
    for _ in range(5):
        pad = np.zeros((h + 4, w + 4), np.uint8)
        pad[2:-2, 2:-2] = cur
        nxt = np.zeros_like(cur)
        for dx, dy in offs:
            nxt = np.maximum(nxt, pad[2 + dy:2 + dy + h, 2 + dx:2 + dx + w])
        cur = nxt
    glare = np.where(cur > 245, 0, 255).astype(np.uint8)
    return glare if internal is None else (np.asarray(internal, np.uint8) & glare)


def shitomasi(img, x, y, half_box=3):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().ora_shitomasi(_p(img), img.shape[1], half_box, int(x), int(y))


def warp_matrix(warp_inv, level):
    """opts::M2Inverse(mm2WarpInverse) * LevelScale(level)  (src/PatchFinder.cc:138)."""
    m = np.asarray(warp_inv, np.float64).reshape(2, 2)
    det = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    idet = 1.0 / det
    r = np.array([[m[1, 1] * idet, -m[0, 1] * idet], [-m[1, 0] * idet, m[0, 0] * idet]])
    return r * float(1 << level)


def patch_template(src_img, m2, cx, cy):
    src_img = np.ascontiguousarray(src_img, np.uint8)
    m2 = np.ascontiguousarray(m2, np.float64)
    t = np.zeros(64, np.uint8)
    nout = lib().ora_patch_template(_p(src_img), src_img.shape[1], src_img.shape[0], src_img.shape[1], _p(m2), float(cx), float(cy), _p(t))
    return t, nout


def find_patch_coarse(img, corners, lut, templ, level, pred, rng, exhaustive=False):
    img = np.ascontiguousarray(img, np.uint8)
    corners = np.ascontiguousarray(corners, np.int32)
    lut = np.ascontiguousarray(lut, np.int32)
    best = np.zeros(2, np.int32)
    score = C.c_int32()
    f = lib().ora_find_patch_coarse(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(corners), len(corners), _p(lut),
                                    _p(np.ascontiguousarray(templ, np.uint8)), level, int(pred[0]), int(pred[1]), rng,
                                    int(exhaustive), _p(best), C.byref(score))
    return bool(f), best, score.value


def subpix(img, templ, level, pos, max_its):
    img = np.ascontiguousarray(img, np.uint8)
    p = np.array(pos, np.float64)
    ok = lib().ora_subpix(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(np.ascontiguousarray(templ, np.uint8)), level, _p(p), max_its)
    return bool(ok), p


def search_patch(pyr_src, pyr_tgt, tgt_levels, req):
    """Oracle of one Tracker::SearchForPoints iteration (src/Tracker.cc:1299-1377) for one request dict/record."""
    res = dict(template_bad=1, found=0, did_subpix=0, score=0, coarse_x=0, coarse_y=0, found_x=0.0, found_y=0.0)
    lvl = int(req["search_level"])
    m2 = warp_matrix(req["warp_inv"], lvl)
    t, nout = patch_template(pyr_src[int(req["src_level"])], m2, req["src_cx"], req["src_cy"])
    res["template"] = t
    if nout:
        return res
    res["template_bad"] = 0
    L = tgt_levels[lvl]
    found, best, score = find_patch_coarse(pyr_tgt[lvl], L["corners"], L["row_lut"], t, lvl, (req["pred_x"], req["pred_y"]),
                                           int(req["range"]), bool(req["exhaustive"]))
    res["score"] = score
    if not found:
        return res
    res.update(found=1, coarse_x=int(best[0]), coarse_y=int(best[1]))
    ls = 1 << lvl
    pos = ((best[0] + 0.5) * ls - 0.5, (best[1] + 0.5) * ls - 0.5)
    res.update(found_x=pos[0], found_y=pos[1])
    if int(req["subpix_its"]) > 0:
        res["did_subpix"] = 1
        ok, p = subpix(pyr_tgt[lvl], t, lvl, pos, int(req["subpix_its"]))
        if not ok:
            res["found"] = 0
        else:
            res.update(found_x=float(p[0]), found_y=float(p[1]))
    return res


def minipatch_find(img_src, img_dst, corners, lut, src_xy, start_xy, rng):
    img_src = np.ascontiguousarray(img_src, np.uint8)
    img_dst = np.ascontiguousarray(img_dst, np.uint8)
    corners = np.ascontiguousarray(corners, np.int32)
    lut = np.ascontiguousarray(lut, np.int32)
    x, y = int(src_xy[0]), int(src_xy[1])
    patch = np.ascontiguousarray(img_src[y - 4:y + 5, x - 4:x + 5]).reshape(-1)
    pos = np.array(start_xy, np.int32)
    f = lib().ora_minipatch_find(_p(img_dst), img_dst.shape[1], img_dst.shape[0], img_dst.shape[1], _p(patch), _p(corners), len(corners),
                                 _p(lut), len(lut), rng, _p(pos))
    return bool(f), pos


def project_point(cam, pose_Rt, pw, right_w, down_w):
    pose_Rt = np.ascontiguousarray(pose_Rt, np.float64); pw = np.ascontiguousarray(pw, np.float64)
    right_w = np.ascontiguousarray(right_w, np.float64); down_w = np.ascontiguousarray(down_w, np.float64)
    px = np.zeros(2); D = np.zeros(4); W = np.zeros(4); vc = np.zeros(3); inim = C.c_int()
    lvl = lib().ora_project_point(C.byref(cam), _p(pose_Rt), _p(pw), _p(right_w), _p(down_w), _p(px), _p(D), _p(W), _p(vc), C.byref(inim))
    return {"px": px, "derivs": D, "warp_inv": W, "v3cam": vc, "in_image": inim.value, "level": lvl}


def search_patches_batch(pyr_src, pyr_tgt, tgt_levels, req):
    """All requests in one C call (CPU timing baseline).  Returns (found flags, positions)."""
    L = lib()
    L.ora_search_patches_batch.argtypes = [C.c_void_p] * 7 + [C.c_int] + [C.c_void_p] * 4
    n = len(req)
    src = [np.ascontiguousarray(x, np.uint8) for x in pyr_src]
    tgt = [np.ascontiguousarray(x, np.uint8) for x in pyr_tgt]
    cor = [np.ascontiguousarray(l["corners"], np.int32) for l in tgt_levels]
    lut = [np.ascontiguousarray(l["row_lut"], np.int32) for l in tgt_levels]
    ptr = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    widths = np.array([x.shape[1] for x in src], np.int32); heights = np.array([x.shape[0] for x in src], np.int32)
    ncor = np.array([len(c) for c in cor], np.int32)
    ri = np.zeros((n, 9), np.int32); m2 = np.zeros((n, 4))
    for i in range(n):
        r = req[i]
        ri[i] = (r["src_level"], r["src_cx"], r["src_cy"], r["search_level"], r["pred_x"], r["pred_y"], r["range"], r["subpix_its"], r["exhaustive"])
        m2[i] = warp_matrix(r["warp_inv"], int(r["search_level"])).reshape(-1)
    xy = np.zeros((n, 2)); flag = np.zeros(n, np.int32)
    L.ora_search_patches_batch(ptr(src), ptr(tgt), _p(widths), _p(heights), ptr(cor), _p(ncor), ptr(lut), n, _p(ri), _p(m2), _p(xy), _p(flag))
    return flag, xy


def calc_jacobian(cam, base_Rt, cfb_Rt, pw):
    L = lib()
    L.ora_calc_jacobian.argtypes = [C.c_void_p] * 7
    b = np.ascontiguousarray(base_Rt, np.float64); c = np.ascontiguousarray(cfb_Rt, np.float64); w = np.ascontiguousarray(pw, np.float64)
    px = np.zeros(2); D = np.zeros(4); J = np.zeros(12)
    inv = L.ora_calc_jacobian(C.byref(cam), _p(b), _p(c), _p(w), _p(px), _p(D), _p(J))
    return px, D, J.reshape(2, 6), inv


def pose_update(found_xy, image_xy, sqrt_inv_noise, jac, found, estimator=0, override_sigma=0.0):
    L = lib()
    L.ora_pose_update.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_double] + [C.c_void_p] * 3
    n = len(found)
    f = np.ascontiguousarray(found_xy, np.float64); im = np.ascontiguousarray(image_xy, np.float64)
    s = np.ascontiguousarray(sqrt_inv_noise, np.float64); j = np.ascontiguousarray(jac, np.float64).reshape(n, 12)
    fl = np.ascontiguousarray(found, np.int32)
    mu = np.zeros(6); sig = C.c_double(); out = np.zeros(n, np.int32)
    nin = L.ora_pose_update(n, _p(f), _p(im), _p(s), _p(j), _p(fl), estimator, float(override_sigma), _p(mu), C.byref(sig), _p(out))
    return mu, sig.value, out, nin
