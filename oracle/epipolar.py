"""CPU restatement of MapMakerServerBase::AddPointEpipolar (src/MapMakerServerBase.cc:604-914) and
ReprojectPoint (:123-143) on top of the oracle's per-patch primitives.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Follows the reference statement by statement, including the behaviour of the ONE PatchFinder / ONE MapPoint object
that the function reuses for every hypothesis: PatchFinder::MakeTemplateCoarseCont (src/PatchFinder.cc:135-182)
only regenerates the template when the warp matrix moved by more than 0.07 since the last generated template, so
consecutive depth hypotheses share templates (and the `template bad` flag).  [3P] TooN SVD<4,4> is numpy.linalg.svd.
"""
import ctypes as C
import math

import numpy as np

from . import oracle as O


def _unproject(cam, px):
    ray = np.zeros(3)
    p = np.ascontiguousarray(px, np.float64)
    O.lib().ora_cam_unproject(C.byref(cam), O._p(p), O._p(ray))
    return ray


def _project(cam, p3):
    px = np.zeros(2)
    D = np.zeros(4)
    p = np.ascontiguousarray(p3, np.float64)
    invalid = O.lib().ora_cam_project(C.byref(cam), O._p(p), O._p(px), O._p(D))
    return bool(invalid), px, D


def one_pixel_angle(cam):
    """TaylorCamera::RefreshParams tail (src/TaylorCamera.cc:192-196)."""
    c = np.array(cam.image_size[:], np.float64) / 2
    a, b = _unproject(cam, c), _unproject(cam, c + 1.0)
    return math.acos(float(a @ b)) / math.sqrt(2.0)


def reproject_point(R_ab, t_ab, vA, vB):
    """MapMakerServerBase::ReprojectPoint (Hartley & Zisserman 12.2): point in frame B seen along vB from B and vA from A."""
    P = np.hstack([R_ab, t_ab.reshape(3, 1)])
    A = np.zeros((4, 4))
    A[0] = (-vB[2], 0.0, vB[0], 0.0)
    A[1] = (0.0, -vB[2], vB[1], 0.0)
    A[2] = vA[0] * P[2] - vA[2] * P[0]
    A[3] = vA[1] * P[2] - vA[2] * P[1]
    v = np.linalg.svd(A)[2][3].copy()
    if v[3] == 0.0:
        v[3] = 0.00001
    return v[:3] / v[3]


def pixel_vectors(src_R, src_t, world, center, right, down, normal=(0.0, 0.0, -1.0)):
    """MapPoint::RefreshPixelVectors (src/MapPoint.cc:62-87)."""
    n = np.asarray(normal, np.float64)
    pc = src_R @ world + src_t
    h = abs(float(pc @ n))
    cp = center * h / abs(float(center @ n))
    rp = right * h / abs(float(right @ n))
    dp = down * h / abs(float(down @ n))
    return src_R.T @ (rp - cp), src_R.T @ (dp - cp)


def hypotheses(cam_tgt, src_R, src_t, tgt_R, tgt_t, ray_sc, level):
    """The depth hypotheses of :620-724: returns None ("return false") or a dict with the list of
    (world position, position in the target camera)."""
    line_tc = tgt_R @ (src_R.T @ ray_sc)
    center_tc = tgt_R @ (-(src_R.T @ src_t)) + tgt_t            # source camera centre in the target frame
    center_sc = src_R @ (-(tgt_R.T @ tgt_t)) + src_t            # target camera centre in the source frame
    max_epi, min_epi = math.pi / 3, 0.05
    sep = float(np.linalg.norm(center_sc))
    src_angle = math.acos(float(center_sc @ ray_sc) / sep)
    start = sep * math.sin(math.pi - src_angle - max_epi) / math.sin(max_epi)
    end = sep * math.sin(math.pi - src_angle - min_epi) / math.sin(min_epi)
    if start < 0.2:
        start = 0.2
    ray_start = center_tc + start * line_tc
    ray_end = center_tc + end * line_tc
    a = ray_start / np.linalg.norm(ray_start)
    b = ray_end / np.linalg.norm(ray_end)
    d = a - b
    if float(d @ d) < 0.00000001:
        return None
    nrm = np.cross(a, b)
    nrm = nrm / np.linalg.norm(nrm)
    pi_, pj = a, np.cross(nrm, a)
    M = np.vstack([pi_, pj, nrm])
    plane_b = (M @ b)[:2]
    max_angle = math.acos(float(plane_b[0]))
    step = one_pixel_angle(cam_tgt) * (1 << level) * 3
    n_steps = int(math.ceil(max_angle / step))
    step = max_angle / n_steps
    s2 = (M @ ray_start)[:2]
    e2 = (M @ ray_end)[:2]
    dir2 = e2 - s2
    dir2 = dir2 / np.linalg.norm(dir2)
    out = []
    for i in range(n_steps + 1):
        ang = i * step
        c = np.array([math.cos(ang), math.sin(ang)])
        alpha = (s2[0] * c[1] - s2[1] * c[0]) / (dir2[1] * c[0] - dir2[0] * c[1])
        p_tc = ray_start + alpha * line_tc
        out.append((tgt_R.T @ (p_tc - tgt_t), p_tc))
    return dict(start_depth=start, end_depth=end, n_steps=n_steps, points=out)


class _Finder:
    """The state of the one PatchFinder the reference function reuses (template cache of MakeTemplateCoarseCont)."""

    def __init__(self, pyr_src, src_level, center):
        self.pyr_src, self.src_level, self.center = pyr_src, src_level, center
        self.last_m2 = None
        self.template = None
        self.template_bad = True
        self.generated = 0

    def make_template(self, warp_inv, level):
        m2 = O.warp_matrix(warp_inv, level)
        refresh = self.last_m2 is None
        if not refresh:
            for i in range(2):
                dv = m2[:, i] - self.last_m2[:, i]
                if float(dv @ dv) > 0.07 * 0.07:
                    refresh = True
        if refresh:
            t, nout = O.patch_template(self.pyr_src[self.src_level], m2, self.center[0], self.center[1])
            self.template, self.template_bad, self.last_m2 = t, bool(nout), m2
            self.generated += 1


def add_point_epipolar(cam_src, cam_tgt, src_cfw, tgt_cfw, pyr_src, pyr_tgt, tgt_levels, level, level_pos, tgt_mask=None):
    """Returns a dict: ok (the reference's return value), and when ok: world (new point), subpix (level-0 position in the
    target keyframe), root (level-0 position in the source keyframe); diagnostics otherwise."""
    src_cfw = np.asarray(src_cfw, np.float64).reshape(-1)
    tgt_cfw = np.asarray(tgt_cfw, np.float64).reshape(-1)
    src_R, src_t = src_cfw[:9].reshape(3, 3), src_cfw[9:]
    tgt_R, tgt_t = tgt_cfw[:9].reshape(3, 3), tgt_cfw[9:]
    ls = 1 << level
    root = np.array([(level_pos[0] + 0.5) * ls - 0.5, (level_pos[1] + 0.5) * ls - 0.5])     # LevelZeroPos
    ray_sc = _unproject(cam_src, root)
    res = dict(ok=False, root=root, reason="")
    hyp = hypotheses(cam_tgt, src_R, src_t, tgt_R, tgt_t, ray_sc, level)
    if hyp is None:
        res["reason"] = "endpoints"
        return res
    res.update(n_steps=hyp["n_steps"], start_depth=hyp["start_depth"], end_depth=hyp["end_depth"])
    center = _unproject(cam_src, root)
    right = _unproject(cam_src, root + np.array([ls, 0.0]))
    down = _unproject(cam_src, root + np.array([0.0, ls]))
    center, right, down = (v / np.linalg.norm(v) for v in (center, right, down))
    finder = _Finder(pyr_src, level, level_pos)
    h0, w0 = pyr_tgt[0].shape
    max_ssd = 8 * 8 * 250 + 1                             # PatchFinder::mnMaxSSD + 1 (src/PatchFinder.cc:44,61)
    best_score, best_i, matches = max_ssd, -1, []

    def prepare(i):
        world, p_tc = hyp["points"][i]
        rw, dw = pixel_vectors(src_R, src_t, world, center, right, down)
        invalid, px, D = _project(cam_tgt, p_tc)
        return world, p_tc, rw, dw, invalid, px, D

    for i in range(len(hyp["points"])):
        world, p_tc, rw, dw, invalid, px, D = prepare(i)
        if invalid:
            continue
        ix, iy = int(px[0]), int(px[1])                   # CVD::ir: truncation
        if not (0 <= ix < w0 and 0 <= iy < h0):
            continue
        if tgt_mask is not None and tgt_mask[iy, ix] == 0:
            continue
        pr = O.project_point(cam_tgt, tgt_cfw, world, rw, dw)
        if pr["level"] == -1:
            continue
        finder.make_template(pr["warp_inv"], pr["level"])
        if finder.template_bad:
            continue
        L = tgt_levels[pr["level"]]
        found, best, score = O.find_patch_coarse(pyr_tgt[pr["level"]], L["corners"], L["row_lut"], finder.template, pr["level"],
                                                 (ix, iy), 3, False)
        if not found:
            continue
        s = 1 << pr["level"]
        coarse = np.array([(best[0] + 0.5) * s - 0.5, (best[1] + 0.5) * s - 0.5])
        matches.append((score, i, coarse))
        if score < best_score:
            best_score, best_i = score, i
    res["n_matches"] = len(matches)
    if best_i == -1:
        res["reason"] = "no match"
        return res
    matches.sort(key=lambda m: m[0])                      # std::sort with compScores (ascending score); ties: see tests
    keep = 1
    for m in matches[1:]:
        if m[0] > best_score * 0.9:                       # as written in the reference (:803)
            keep += 1
    if keep > 3:
        res["reason"] = "ambiguous count"
        return res
    matches = matches[:keep]
    for m in matches[1:]:
        if abs(m[1] - best_i) > 1:
            res["reason"] = "ambiguous index"
            return res
    subpix = None
    for score, i, coarse in matches:
        world, p_tc, rw, dw, invalid, px, D = prepare(i)
        pr = O.project_point(cam_tgt, tgt_cfw, world, rw, dw)
        finder.make_template(pr["warp_inv"], pr["level"])
        ok, p = O.subpix(pyr_tgt[pr["level"]], finder.template, pr["level"], coarse, 10)
        if not ok:
            continue
        subpix = p
        res["subpix_from"] = i
        break
    if subpix is None:
        res["reason"] = "subpix"
        return res
    # A = source camera, B = target camera: se3AfromB = src * tgt^-1
    R_ab = src_R @ tgt_R.T
    t_ab = src_t - R_ab @ tgt_t
    p_b = reproject_point(R_ab, t_ab, _unproject(cam_src, root), _unproject(cam_tgt, subpix))
    res.update(ok=True, subpix=subpix, world=tgt_R.T @ (p_b - tgt_t), best=best_i, best_score=best_score,
               templates_generated=finder.generated)
    return res
