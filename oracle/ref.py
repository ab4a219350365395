"""ctypes loader for oracle/_ref/libref.so: the REFERENCE'S OWN translation units (MEstimator.h, ShiTomasi.cc, MiniPatch.cc,
TaylorCamera.cc, PatchFinder.cc) compiled from /root/reference by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref

_libs = {}


def lib(scalar: bool = False):
    """None when neither the reference tree nor a prebuilt library is present."""
    key = "scalar" if scalar else "sse"
    if key not in _libs:
        path = build_ref.build()
        if path is None:
            _libs[key] = None
        else:
            if scalar:
                path = os.path.join(os.path.dirname(path), "libref_scalar.so")
            L = C.CDLL(path)
            for f in ("ref_huber_sigma_sq", "ref_tukey_sigma_sq", "ref_cauchy_sigma_sq", "ref_shitomasi", "ref_level_zero_pos", "ref_level_n_pos"):
                getattr(L, f).restype = C.c_double
            L.ref_huber_sigma_sq.argtypes = L.ref_tukey_sigma_sq.argtypes = L.ref_cauchy_sigma_sq.argtypes = [C.c_void_p, C.c_int]
            L.ref_mestimator_weights.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_shitomasi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            L.ref_minipatch_find.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_level_zero_pos.argtypes = L.ref_level_n_pos.argtypes = [C.c_double, C.c_int]
            L.ref_cam_create.restype = C.c_void_p
            L.ref_cam_create.argtypes = [C.c_void_p] + [C.c_int] * 6
            L.ref_cam_destroy.argtypes = [C.c_void_p]
            L.ref_cam_derived.argtypes = [C.c_void_p] * 8
            L.ref_cam_project.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
            L.ref_cam_unproject.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.ref_zmssd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
            L.ref_calc_search_level.argtypes = [C.c_void_p] * 6
            L.ref_patch_search.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int] +
                                           [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 3)
            _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefCamera:
    """TaylorCamera(v9Params, calib, full, image) of the reference (src/TaylorCamera.cc)."""

    def __init__(self, params9, calib=(640, 480), full=(640, 480), image=(640, 480)):
        self.L = lib()
        p = np.ascontiguousarray(params9, np.float64)
        self.h = self.L.ref_cam_create(_p(p), calib[0], calib[1], full[0], full[1], image[0], image[1])

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_cam_destroy(self.h)
            self.h = None

    def derived(self):
        center = np.zeros(2); affine = np.zeros(4); inv = np.zeros(32)
        mt, mean, std, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        self.L.ref_cam_derived(self.h, _p(center), _p(affine), C.byref(mt), C.byref(mean), C.byref(std), C.byref(n), _p(inv))
        return dict(center=center, affine=affine, min_theta=mt.value, theta_mean=mean.value, theta_std=std.value, inv_poly=inv[:n.value].copy())

    def project(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = len(xyz)
        px = np.zeros((n, 2)); inv = np.zeros(n, np.int32); D = np.zeros((n, 4)); dt = np.zeros((n, 3)); dp = np.zeros((n, 3))
        self.L.ref_cam_project(self.h, n, _p(xyz), _p(px), _p(inv), _p(D), _p(dt), _p(dp))
        return px, inv, D, dt, dp

    def unproject(self, px):
        px = np.ascontiguousarray(px, np.float64)
        ray = np.zeros((len(px), 3))
        self.L.ref_cam_unproject(self.h, len(px), _p(px), _p(ray))
        return ray


def patch_search(src_img, req, tgt_img, corners, row_lut, scalar=False):
    """One Tracker::SearchForPoints iteration on the reference's PatchFinder (source level 0 image `src_img`, target level image)."""
    L = lib(scalar)
    src_img = np.ascontiguousarray(src_img, np.uint8); tgt_img = np.ascontiguousarray(tgt_img, np.uint8)
    corners = np.ascontiguousarray(corners, np.int32); row_lut = np.ascontiguousarray(row_lut, np.int32)
    w = np.ascontiguousarray(req["warp_inv"], np.float64)
    out6 = np.zeros(6, np.int32); xy = np.zeros(2); t = np.zeros(64, np.uint8)
    L.ref_patch_search(_p(src_img), src_img.shape[1], src_img.shape[0], src_img.shape[1], int(req["src_cx"]), int(req["src_cy"]), _p(w), int(req["search_level"]),
                       _p(tgt_img), tgt_img.shape[1], tgt_img.shape[0], tgt_img.shape[1], _p(corners), len(corners), _p(row_lut),
                       int(req["pred_x"]), int(req["pred_y"]), int(req["range"]), int(req["subpix_its"]), int(req["exhaustive"]), _p(out6), _p(xy), _p(t))
    return dict(template_bad=int(out6[0]), found=int(out6[1]), did_subpix=int(out6[2]), score=int(out6[3]), coarse_x=int(out6[4]), coarse_y=int(out6[5]),
                found_x=float(xy[0]), found_y=float(xy[1]), template=t)


# ---------------------------------------------------------------------------------------------------------------------
# src/ChainBundle.cc (compiled unmodified against the g2o stand-in): libref_ba.so
# ---------------------------------------------------------------------------------------------------------------------
_ba_lib = None


def ba_lib():
    global _ba_lib
    if _ba_lib is None:
        path = build_ref.build()
        if path is None:
            return None
        L = C.CDLL(os.path.join(os.path.dirname(path), "libref_ba.so"))
        L.ref_ba_create.restype = C.c_void_p
        L.ref_ba_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_ba_destroy.argtypes = [C.c_void_p]
        L.ref_ba_load.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5
        L.ref_ba_eval.argtypes = [C.c_void_p] * 3
        L.ref_ba_jacobians.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        L.ref_ba_oplus_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_ba_oplus_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_ba_robustify.argtypes = [C.c_void_p] * 3
        L.ref_ba_compute.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.ref_ba_get_state.argtypes = [C.c_void_p] * 3
        L.ref_ba_get_outliers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _ba_lib = L
    return _ba_lib


class RefBA:
    """The reference's ChainBundle on a flat problem (synth.BaProblem).  `cam_params` = the 9 Taylor parameters per camera
    the problem's camera structs were derived from (the reference derives everything else itself, RefreshParams)."""

    def __init__(self, prob, cam_params, use_robust=True, use_tukey=True, image=(640, 480)):
        self.L = ba_lib()
        self.prob = prob
        p9 = np.ascontiguousarray(cam_params, np.float64).reshape(-1, 9)
        sizes = np.ascontiguousarray(np.tile(np.array([image[0], image[1]] * 3, np.int32), (len(p9), 1)))
        self.h = self.L.ref_ba_create(len(p9), _p(p9), _p(sizes), int(use_robust), int(use_tukey))
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        self._keep = [a(prob.pose_Rt, np.float64), a(prob.pose_fixed, np.uint8), a(prob.pt_xyz, np.float64), a(prob.pt_chain, np.int32), a(prob.pt_fixed, np.uint8),
                      a(prob.meas_xy, np.float64), a(prob.meas_chain, np.int32), a(prob.meas_pt, np.int32), a(prob.meas_noise, np.float64), a(prob.meas_cam, np.int32)]
        k = self._keep
        rc = self.L.ref_ba_load(self.h, prob.n_pose, _p(k[0]), _p(k[1]), prob.n_pt, _p(k[2]), _p(k[3]), _p(k[4]), prob.n_meas, _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]))
        assert rc == 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_ba_destroy(self.h)
            self.h = None

    def eval(self):
        e = np.zeros((self.prob.n_meas, 2)); c = np.zeros(self.prob.n_meas)
        self.L.ref_ba_eval(self.h, _p(e), _p(c))
        return e, c

    def jacobians(self, m):
        jo = np.zeros((2, 2, 6)); js = np.zeros((2, 2, 6)); jp = np.zeros((2, 3))
        self.L.ref_ba_jacobians(self.h, int(m), _p(jo), _p(js), _p(jp))
        return jo, js, jp

    def oplus_pose(self, i, d):
        out = np.zeros(12); d = np.ascontiguousarray(d, np.float64)
        self.L.ref_ba_oplus_pose(self.h, int(i), _p(d), _p(out))
        return out

    def oplus_point(self, p, d):
        out = np.zeros(3); d = np.ascontiguousarray(d, np.float64)
        self.L.ref_ba_oplus_point(self.h, int(p), _p(d), _p(out))
        return out

    def robustify(self):
        rho = np.zeros((self.prob.n_meas, 3)); s = C.c_double()
        self.L.ref_ba_robustify(self.h, _p(rho), C.byref(s))
        return rho, s.value

    def compute(self, n_iter, user_lambda=-1.0):
        st = np.zeros(6)
        rc = self.L.ref_ba_compute(self.h, int(n_iter), float(user_lambda), _p(st))
        return rc, dict(total_trials=int(st[0]), converged=int(st[1]), sigma_sq=st[2], mean_chi2=st[3], lambda_=st[4], max_cov=st[5])

    def state(self):
        P = np.zeros((self.prob.n_pose, 12)); X = np.zeros((self.prob.n_pt, 3))
        self.L.ref_ba_get_state(self.h, _p(P), _p(X))
        return P, X

    def outliers(self):
        idx = np.zeros(max(self.prob.n_meas, 1), np.int32)
        n = self.L.ref_ba_get_outliers(self.h, _p(self._keep[6]), _p(self._keep[9]), _p(idx), len(idx))
        return np.sort(idx[:n])
