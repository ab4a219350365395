"""ctypes plumbing over the C ABI (include/mcptam_b200.h) for tests and bench.py.

This is not the product: the product is libmcptam_b200.so (CUDA kernels + extern "C" layer) and the C++
host mirror under mcptam_b200/host/.  There is no CPU fallback: if the library is missing or no GPU is
present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .synth import TaylorCamStruct

MCP_OK = 0
ERRORS = {-101: "MCP_ERR_INVALID", -102: "MCP_ERR_CUDA", -103: "MCP_ERR_UNSUPPORTED", -104: "MCP_ERR_NO_DEVICE",
          -105: "MCP_ERR_NCCL", -106: "MCP_ERR_STATE"}


class McpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "error"), code, msg))
        self.code = code


class BaConfig(C.Structure):
    _fields_ = [("use_robust", C.c_int32), ("use_tukey", C.c_int32), ("verbose", C.c_int32),
                ("max_trials_after_failure", C.c_int32), ("update_pct_limit", C.c_double),
                ("update_rms_limit", C.c_double), ("min_sigma", C.c_double), ("device", C.c_int32), ("pad_", C.c_int32)]


class BaStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("total_trials", C.c_int32), ("converged", C.c_int32),
                ("hit_max_iter", C.c_int32), ("n_outliers", C.c_int32), ("pad_", C.c_int32),
                ("sigma_sq", C.c_double), ("mean_chi2", C.c_double), ("lambda_", C.c_double), ("max_cov", C.c_double),
                ("chi2_before", C.c_double), ("chi2_after", C.c_double), ("gpu_ms", C.c_double),
                ("kernel_launches", C.c_int32), ("pad2_", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("pad")}


class BaTiming(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_select", "ms_linearize", "ms_schur", "ms_solve", "ms_backsub", "ms_control", "ms_other")] + \
               [(n, C.c_int32) for n in ("n_select", "n_linearize", "n_schur", "n_solve", "n_backsub", "n_control", "n_other", "pad_")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("pad")}


_lib = None


def lib_path() -> str:
    return _build.LIB


def _preload_bundled_nccl():
    """torch ships its own libnccl.so.2 under site-packages/nvidia/nccl -- the same SONAME as the system NCCL the library
    is linked against, so whichever is loaded first serves both.  Loading the bundled one first keeps a later
    `import torch` in the same process working (its libtorch_cuda.so needs symbols the older system NCCL lacks)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass                                    # no bundled NCCL: the system one is used


def lib():
    """Loads libmcptam_b200.so (building it with nvcc if the tree is newer)."""
    global _lib
    if _lib is None:
        path = _build.build()
        _preload_bundled_nccl()
        if not os.path.exists(path):
            raise RuntimeError("libmcptam_b200.so is missing: the CUDA extension must be built (no CPU fallback)")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.mcp_last_error.restype = C.c_char_p
        L.mcp_ba_default_config.argtypes = [C.POINTER(BaConfig)]
        L.mcp_ba_create.argtypes = [C.POINTER(BaConfig), C.POINTER(C.c_void_p)]
        L.mcp_ba_destroy.argtypes = [C.c_void_p]
        L.mcp_ba_set_cameras.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.mcp_ba_load.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcp_ba_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.POINTER(BaStats)]
        L.mcp_ba_get_poses.argtypes = [C.c_void_p, C.c_void_p]
        L.mcp_ba_get_points.argtypes = [C.c_void_p, C.c_void_p]
        L.mcp_ba_get_outliers.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.mcp_ba_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcp_ba_reset_state.argtypes = [C.c_void_p]
        L.mcp_ba_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.mcp_nccl_unique_id.argtypes = [C.c_void_p]
        L.mcp_ba_partition.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.mcp_ba_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcp_ba_debug_jacobians.argtypes = [C.c_void_p, C.c_void_p]
        L.mcp_ba_lm_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcp_ba_debug_solve_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.mcp_ba_get_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.mcp_ba_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        L.mcp_ba_get_timing.argtypes = [C.c_void_p, C.POINTER(BaTiming)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def check(rc):
    if rc < 0:
        raise McpError(rc, lib().mcp_last_error().decode())
    return rc


def cam_array(cams):
    arr = (TaylorCamStruct * len(cams))()
    for i, c in enumerate(cams):
        C.memmove(C.byref(arr[i]), C.byref(c), C.sizeof(c))
    return arr


class BaHandle:
    """One mcp_ba handle: load a BaProblem (mcptam_b200.synth.BaProblem-like) and run Compute."""

    def __init__(self, use_robust=True, use_tukey=True, device=-1, **cfg_kw):
        self.L = lib()
        cfg = BaConfig()
        self.L.mcp_ba_default_config(C.byref(cfg))
        cfg.use_robust = int(use_robust)
        cfg.use_tukey = int(use_tukey)
        cfg.device = device
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        self.h = C.c_void_p()
        check(self.L.mcp_ba_create(C.byref(cfg), C.byref(self.h)))
        self.prob = None

    def close(self):
        if self.h:
            self.L.mcp_ba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id: bytes | None, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self.L.mcp_ba_comm_init(self.h, buf, rank, world))

    def load(self, prob):
        self.prob = prob
        self._cams = cam_array(prob.cams)
        check(self.L.mcp_ba_set_cameras(self.h, len(prob.cams), C.cast(self._cams, C.c_void_p)))
        k = [np.ascontiguousarray(prob.pose_Rt, np.float64), np.ascontiguousarray(prob.pose_fixed, np.uint8),
             np.ascontiguousarray(prob.pt_xyz, np.float64), np.ascontiguousarray(prob.pt_chain, np.int32),
             np.ascontiguousarray(prob.pt_fixed, np.uint8), np.ascontiguousarray(prob.meas_xy, np.float64),
             np.ascontiguousarray(prob.meas_chain, np.int32), np.ascontiguousarray(prob.meas_pt, np.int32),
             np.ascontiguousarray(prob.meas_noise, np.float64), np.ascontiguousarray(prob.meas_cam, np.int32)]
        self._keep = k
        check(self.L.mcp_ba_load(self.h, prob.n_pose, _p(k[0]), _p(k[1]), prob.n_pt, _p(k[2]), _p(k[3]), _p(k[4]),
                                 prob.n_meas, _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9])))
        self.n_pose_var = int((k[1] == 0).sum())
        self.n_pt_var = int((k[4] == 0).sum())

    def compute(self, n_iter=100, user_lambda=-1.0, abort=None):
        st = BaStats()
        rc = self.L.mcp_ba_compute(self.h, _p(abort) if abort is not None else None, n_iter, float(user_lambda), C.byref(st))
        if rc < -1:
            check(rc)
        return rc, st

    def poses(self):
        o = np.zeros((self.prob.n_pose, 12))
        check(self.L.mcp_ba_get_poses(self.h, _p(o)))
        return o

    def points(self):
        o = np.zeros((self.prob.n_pt, 3))
        check(self.L.mcp_ba_get_points(self.h, _p(o)))
        return o

    def outliers(self):
        n = self.L.mcp_ba_get_outliers(self.h, None, 0)
        o = np.zeros(max(n, 1), np.int32)
        self.L.mcp_ba_get_outliers(self.h, _p(o), n)
        return o[:n]

    def set_state(self, poses, points):
        poses = np.ascontiguousarray(poses, np.float64)
        points = np.ascontiguousarray(points, np.float64)
        check(self.L.mcp_ba_set_state(self.h, _p(poses), _p(points)))

    def reset_state(self):
        check(self.L.mcp_ba_reset_state(self.h))

    def eval(self):
        e = np.zeros((self.prob.n_meas, 2))
        c = np.zeros(self.prob.n_meas)
        check(self.L.mcp_ba_eval(self.h, _p(e), _p(c)))
        return e, c

    def jacobians(self):
        j = np.zeros((self.prob.n_meas, 30))
        check(self.L.mcp_ba_debug_jacobians(self.h, _p(j)))
        return j

    def lm_step(self, lam, sigma_sq=-1.0):
        d = np.zeros(6 * self.n_pose_var + 3 * self.n_pt_var)
        s = C.c_double()
        r = C.c_double()
        check(self.L.mcp_ba_lm_step(self.h, float(lam), float(sigma_sq), _p(d), C.byref(s), C.byref(r)))
        return d, s.value, r.value

    def solve_trace(self, arm=False):
        if arm:
            check(self.L.mcp_ba_debug_solve_trace(self.h, None, 0))
            return None
        o = np.zeros(8 * 2048)
        check(self.L.mcp_ba_debug_solve_trace(self.h, _p(o), len(o)))
        return o.reshape(-1, 8)

    def stream(self) -> int:
        p = C.c_void_p()
        check(self.L.mcp_ba_get_stream(self.h, C.byref(p)))
        return int(p.value or 0)

    def set_profiling(self, on=True):
        check(self.L.mcp_ba_set_profiling(self.h, int(on)))

    def timing(self):
        t = BaTiming()
        check(self.L.mcp_ba_get_timing(self.h, C.byref(t)))
        return t.as_dict()


def ba_partition(n_pt, meas_pt, world):
    meas_pt = np.ascontiguousarray(meas_pt, np.int32)
    out = np.zeros(world + 1, np.int32)
    check(lib().mcp_ba_partition(n_pt, _p(meas_pt), len(meas_pt), world, _p(out)))
    return out


class _PrepView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pose_var pt_var pt_info pt_order pt_meas_off pt_slot_off slot_var slot_pt meas_xy meas_info "
                                          "meas_a meas_b pb_idx pb_items rs_ent rs_grp rs_items meas_orig part_pt part_meas").split()] + \
               [(n, C.c_longlong) for n in "n_pb_idx n_pb_items n_rs_ent n_rs_grp n_rs_items n_inc".split()] + \
               [(n, C.c_int) for n in "npv nptv n_slots max_slots rs_nblk pad".split()] + [("err", C.c_char_p)]


def ba_prepare(prob, rank=0, world=1, want_rows=False, reps=1, threads=1, par_min_meas=-1):
    """Host marshalling of mcp_ba_load (csrc/ba_prep.hpp) run on the CPU through the test shim libmcptam_prep.so.
    Returns (dict of numpy arrays / scalars, best time in ms).  Raises McpError with the message mcp_ba_load would set."""
    from . import build
    if not os.path.exists(build.PREP_LIB):
        build.build(force=True)
    L = C.CDLL(build.PREP_LIB)
    k = [np.ascontiguousarray(prob.pose_fixed, np.uint8), np.ascontiguousarray(prob.pt_chain, np.int32),
         np.ascontiguousarray(prob.pt_fixed, np.uint8), np.ascontiguousarray(prob.meas_xy, np.float64),
         np.ascontiguousarray(prob.meas_chain, np.int32), np.ascontiguousarray(prob.meas_pt, np.int32),
         np.ascontiguousarray(prob.meas_noise, np.float64), np.ascontiguousarray(prob.meas_cam, np.int32)]
    v, ms = _PrepView(), C.c_double()
    L.mcp_prep_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + \
                              [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
    rc = L.mcp_prep_run(len(prob.cams), prob.n_pose, _p(k[0]), prob.n_pt, _p(k[1]), _p(k[2]), prob.n_meas, _p(k[3]), _p(k[4]),
                        _p(k[5]), _p(k[6]), _p(k[7]), rank, world, int(want_rows), reps, threads, par_min_meas, C.cast(C.byref(ms), C.c_void_p),
                        C.cast(C.byref(v), C.c_void_p))
    if rc != 0:
        raise McpError(rc, (v.err or b"").decode())

    def arr(ptr, n, dt=np.int32, cols=1):
        if not ptr or n == 0:
            return np.zeros((0, cols) if cols > 1 else 0, dt)
        ct = C.c_double if dt == np.float64 else C.c_int
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (int(n) * cols,)).copy()
        return a.reshape(-1, cols) if cols > 1 else a

    n_pt, n_meas = prob.n_pt, prob.n_meas
    out = dict(pose_var=arr(v.pose_var, prob.n_pose), pt_var=arr(v.pt_var, n_pt), pt_info=arr(v.pt_info, n_pt, cols=4),
               pt_order=arr(v.pt_order, n_pt), pt_meas_off=arr(v.pt_meas_off, n_pt + 1), pt_slot_off=arr(v.pt_slot_off, n_pt + 1),
               slot_var=arr(v.slot_var, v.n_slots), slot_pt=arr(v.slot_pt, v.n_slots), meas_xy=arr(v.meas_xy, n_meas, np.float64, 2),
               meas_info=arr(v.meas_info, n_meas, np.float64), meas_a=arr(v.meas_a, n_meas, cols=4), meas_b=arr(v.meas_b, n_meas, cols=4),
               pb_idx=arr(v.pb_idx, v.n_pb_idx), pb_items=arr(v.pb_items, v.n_pb_items, cols=4), rs_ent=arr(v.rs_ent, v.n_rs_ent, cols=2),
               rs_grp=arr(v.rs_grp, v.n_rs_grp), rs_items=arr(v.rs_items, v.n_rs_items, cols=4), meas_orig=arr(v.meas_orig, n_meas),
               part_pt=arr(v.part_pt, world + 1), part_meas=arr(v.part_meas, world + 1), n_inc=int(v.n_inc), npv=v.npv, nptv=v.nptv,
               n_slots=v.n_slots, max_slots=v.max_slots, rs_nblk=v.rs_nblk)
    return out, ms.value


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().mcp_nccl_unique_id(buf))
    return buf.raw


# ---------------------------------------------------------------------------------------------
# front end
# ---------------------------------------------------------------------------------------------
class FeConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("adaptive_thresh", C.c_int32),
                ("max_corners_per_level", C.c_int32), ("max_keyframes", C.c_int32), ("max_patches", C.c_int32),
                ("device", C.c_int32), ("halfsample_round", C.c_int32), ("transform_round", C.c_int32), ("pad_", C.c_int32)]


class LevelOut(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_corners", C.c_int32), ("fast_thresh", C.c_int32),
                ("fast_freq", C.c_int32 * 31), ("n_corners_total", C.c_int32), ("image", C.c_void_p), ("corners_xy", C.c_void_p),
                ("corners_cap", C.c_int32), ("pad2_", C.c_int32), ("row_lut", C.c_void_p), ("last_mask", C.c_void_p)]


class PatchReq(C.Structure):
    _fields_ = [("src_kf", C.c_int32), ("src_level", C.c_int32), ("src_cx", C.c_int32), ("src_cy", C.c_int32),
                ("warp_inv", C.c_double * 4), ("search_level", C.c_int32), ("pred_x", C.c_int32), ("pred_y", C.c_int32),
                ("range", C.c_int32), ("subpix_its", C.c_int32), ("exhaustive", C.c_int32)]


class PatchRes(C.Structure):
    _fields_ = [("template_bad", C.c_int32), ("found", C.c_int32), ("did_subpix", C.c_int32), ("score", C.c_int32),
                ("coarse_x", C.c_int32), ("coarse_y", C.c_int32), ("found_x", C.c_double), ("found_y", C.c_double),
                ("n_candidates", C.c_int32), ("pad_", C.c_int32)]


class FeTiming(C.Structure):
    _fields_ = [("ms_pyramid", C.c_double), ("ms_fast", C.c_double), ("ms_compact", C.c_double), ("ms_search", C.c_double),
                ("ms_other", C.c_double), ("n_launches", C.c_int32), ("pad_", C.c_int32)]


class RestConfig(C.Structure):
    _fields_ = [("use_shi", C.c_int32), ("use_thresh", C.c_int32), ("top_fraction", C.c_double), ("thresh", C.c_double),
                ("nonmax_strict", C.c_int32), ("prev_slot", C.c_int32), ("n_prev", C.c_int32), ("pad_", C.c_int32)]


class RestLevelOut(C.Structure):
    _fields_ = [("n_max", C.c_int32), ("n_selected", C.c_int32), ("n_candidates", C.c_int32), ("cap", C.c_int32), ("cand", C.c_void_p)]


CANDIDATE_DTYPE = np.dtype([("x", "i4"), ("y", "i4"), ("score", "f8")], align=True)
PATCH_REQ_DTYPE = np.dtype([("src_kf", "i4"), ("src_level", "i4"), ("src_cx", "i4"), ("src_cy", "i4"), ("warp_inv", "f8", 4),
                            ("search_level", "i4"), ("pred_x", "i4"), ("pred_y", "i4"), ("range", "i4"), ("subpix_its", "i4"),
                            ("exhaustive", "i4")], align=True)
JAC_RES_DTYPE = np.dtype([("px", "f8", 2), ("jac", "f8", 12), ("in_image", "i4"), ("pad_", "i4")], align=True)
POSE_MEAS_DTYPE = np.dtype([("found", "f8", 2), ("image", "f8", 2), ("sqrt_inv_noise", "f8"), ("jac", "f8", 12), ("found_flag", "i4"), ("pad_", "i4")], align=True)
POSE_UPDATE_DTYPE = np.dtype([("mu", "f8", 6), ("sigma_sq", "f8"), ("c_inv", "f8", 36), ("n_inliers", "i4"), ("n_valid", "i4")], align=True)
PROJ_RES_DTYPE = np.dtype([("px", "f8", 2), ("cam_derivs", "f8", 4), ("warp_inv", "f8", 4), ("v3cam", "f8", 3), ("in_image", "i4"), ("search_level", "i4")], align=True)
PATCH_RES_DTYPE = np.dtype([("template_bad", "i4"), ("found", "i4"), ("did_subpix", "i4"), ("score", "i4"), ("coarse_x", "i4"),
                            ("coarse_y", "i4"), ("found_x", "f8"), ("found_y", "f8"), ("n_candidates", "i4"), ("pad_", "i4")], align=True)
assert PATCH_REQ_DTYPE.itemsize == C.sizeof(PatchReq) and PATCH_RES_DTYPE.itemsize == C.sizeof(PatchRes)

_fe_bound = False


def _bind_fe(L):
    global _fe_bound
    if _fe_bound:
        return
    L.mcp_fe_default_config.argtypes = [C.POINTER(FeConfig)]
    L.mcp_fe_create.argtypes = [C.POINTER(FeConfig), C.POINTER(C.c_void_p)]
    L.mcp_fe_destroy.argtypes = [C.c_void_p]
    L.mcp_fe_set_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.mcp_fe_set_glare_masking.argtypes = [C.c_void_p, C.c_int32]
    L.mcp_fe_make_keyframe.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    L.mcp_fe_search_patches.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.mcp_fe_get_templates.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.mcp_fe_shitomasi.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.mcp_fe_minipatch_find.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p]
    L.mcp_fe_get_timing.argtypes = [C.c_void_p, C.POINTER(FeTiming)]
    L.mcp_fe_set_camera.argtypes = [C.c_void_p, C.c_void_p]
    L.mcp_fe_calc_jacobians.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.mcp_fe_pose_update.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
    L.mcp_fe_project_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcp_fe_debug_scores.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.mcp_fe_default_rest_config.argtypes = [C.POINTER(RestConfig)]
    L.mcp_fe_default_rest_config.restype = None
    L.mcp_fe_make_keyframe_rest.argtypes = [C.c_void_p, C.c_int32, C.POINTER(RestConfig), C.c_void_p]
    _fe_bound = True


class FeHandle:
    """One mcp_fe handle = one camera: resident keyframe pyramids + corner lists on the device."""

    def __init__(self, width=640, height=480, device=-1, **cfg_kw):
        self.L = lib()
        _bind_fe(self.L)
        cfg = FeConfig()
        self.L.mcp_fe_default_config(C.byref(cfg))
        cfg.width, cfg.height, cfg.device = width, height, device
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = C.c_void_p()
        check(self.L.mcp_fe_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.mcp_fe_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mask(self, mask):
        if mask is None:
            check(self.L.mcp_fe_set_mask(self.h, None, 0))
        else:
            m = np.ascontiguousarray(mask, np.uint8)
            check(self.L.mcp_fe_set_mask(self.h, _p(m), m.shape[1]))

    def set_glare_masking(self, on=True):
        check(self.L.mcp_fe_set_glare_masking(self.h, int(bool(on))))

    def make_keyframe(self, slot, img, want_images=False, outputs=True, want_masks=False):
        """MakeKeyFrame_Lite of one image into resident slot `slot`.  Returns a list of 4 per-level dicts."""
        img = np.ascontiguousarray(img, np.uint8)
        if not outputs:
            check(self.L.mcp_fe_make_keyframe(self.h, slot, _p(img), img.shape[1], None))
            return None
        cap = self.cfg.max_corners_per_level
        w, h = self.cfg.width, self.cfg.height
        st = getattr(self, "_kf_state", None)
        if st is None:
            # the McpLevelOut array, the corner / LUT landing buffers and a numpy view of the scalar fields are made once per
            # handle: a call then costs the C call plus the copies of what it returns
            outs = (LevelOut * 4)()
            bufs = [(np.zeros((cap, 2), np.int32), np.zeros(h >> l, np.int32)) for l in range(4)]
            for l in range(4):
                outs[l].corners_xy = bufs[l][0].ctypes.data
                outs[l].corners_cap = cap
                outs[l].row_lut = bufs[l][1].ctypes.data
            ints = np.frombuffer(outs, dtype=np.int32).reshape(4, C.sizeof(LevelOut) // 4)
            f = {k: getattr(LevelOut, k).offset // 4 for k in ("width", "height", "n_corners", "fast_thresh", "fast_freq", "n_corners_total")}
            st = self._kf_state = (outs, bufs, ints, f, C.cast(outs, C.c_void_p))
        outs, bufs, ints, f, outs_p = st
        keep = []
        for l in range(4):
            im = np.zeros((h >> l, w >> l), np.uint8) if want_images else None
            mk = np.zeros((h >> l, w >> l), np.uint8) if want_masks else None
            keep.append((im, mk))
            outs[l].image = im.ctypes.data if im is not None else None
            outs[l].last_mask = mk.ctypes.data if mk is not None else None
        check(self.L.mcp_fe_make_keyframe(self.h, slot, _p(img), img.shape[1], outs_p))
        res = []
        for l in range(4):
            cor, lut = bufs[l]
            im, mk = keep[l]
            row = ints[l]
            n = int(row[f["n_corners"]])
            res.append({"width": int(row[f["width"]]), "height": int(row[f["height"]]), "n_corners": n, "corners": cor[:n].copy(),
                        "row_lut": lut.copy(), "n_corners_total": int(row[f["n_corners_total"]]), "fast_thresh": int(row[f["fast_thresh"]]),
                        "fast_freq": row[f["fast_freq"]: f["fast_freq"] + 31].copy(), "image": im, "last_mask": mk})
        return res

    def search_patches(self, target_kf, req: np.ndarray) -> np.ndarray:
        req = np.ascontiguousarray(req, PATCH_REQ_DTYPE)
        res = np.zeros(len(req), PATCH_RES_DTYPE)
        check(self.L.mcp_fe_search_patches(self.h, target_kf, len(req), _p(req), _p(res)))
        return res

    def templates(self, n):
        t = np.zeros((n, 64), np.uint8)
        check(self.L.mcp_fe_get_templates(self.h, n, _p(t)))
        return t

    def shitomasi(self, kf, level, xy):
        xy = np.ascontiguousarray(xy, np.int32)
        out = np.zeros(len(xy))
        check(self.L.mcp_fe_shitomasi(self.h, kf, level, len(xy), _p(xy), _p(out)))
        return out

    def minipatch_find(self, kf_src, kf_dst, level, src_xy, start_xy, rng):
        src_xy = np.ascontiguousarray(src_xy, np.int32)
        start_xy = np.ascontiguousarray(start_xy, np.int32)
        pos = np.zeros_like(src_xy)
        found = np.zeros(len(src_xy), np.int32)
        check(self.L.mcp_fe_minipatch_find(self.h, kf_src, kf_dst, level, len(src_xy), _p(src_xy), _p(start_xy), rng, _p(pos), _p(found)))
        return pos, found

    def make_keyframe_rest(self, slot, prev_slot=-1, n_prev=0, **cfg_kw):
        """KeyFrame::MakeKeyFrame_Rest candidate generation for the pyramid in `slot`; returns 4 per-level dicts."""
        cfg = RestConfig()
        self.L.mcp_fe_default_rest_config(C.byref(cfg))
        cfg.prev_slot, cfg.n_prev = prev_slot, n_prev
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        outs = (RestLevelOut * 4)()
        cap = self.cfg.max_corners_per_level
        keep = []
        for l in range(4):
            c = np.zeros(cap, CANDIDATE_DTYPE)
            keep.append(c)
            outs[l].cand = c.ctypes.data
            outs[l].cap = cap
        check(self.L.mcp_fe_make_keyframe_rest(self.h, slot, C.byref(cfg), C.cast(outs, C.c_void_p)))
        return [{"n_max": outs[l].n_max, "n_selected": outs[l].n_selected, "n_candidates": outs[l].n_candidates,
                 "cand": keep[l][:outs[l].n_candidates].copy()} for l in range(4)]

    def set_camera(self, cam):
        self._cam = cam
        check(self.L.mcp_fe_set_camera(self.h, C.byref(cam)))

    def project_points(self, cam_from_world, world_xyz, right_w, down_w):
        T = np.ascontiguousarray(cam_from_world, np.float64)
        pw = np.ascontiguousarray(world_xyz, np.float64); rw = np.ascontiguousarray(right_w, np.float64); dw = np.ascontiguousarray(down_w, np.float64)
        out = np.zeros(len(pw), PROJ_RES_DTYPE)
        check(self.L.mcp_fe_project_points(self.h, _p(T), len(pw), _p(pw), _p(rw), _p(dw), _p(out)))
        return out

    def calc_jacobians(self, base_from_world, cam_from_base, world_xyz):
        b = np.ascontiguousarray(base_from_world, np.float64); cb = np.ascontiguousarray(cam_from_base, np.float64)
        pw = np.ascontiguousarray(world_xyz, np.float64)
        out = np.zeros(len(pw), JAC_RES_DTYPE)
        check(self.L.mcp_fe_calc_jacobians(self.h, _p(b), _p(cb), len(pw), _p(pw), _p(out)))
        return out

    def pose_update(self, meas, estimator=0, override_sigma=0.0):
        meas = np.ascontiguousarray(meas, POSE_MEAS_DTYPE)
        out = np.zeros(1, POSE_UPDATE_DTYPE)
        outlier = np.zeros(max(len(meas), 1), np.int32)
        check(self.L.mcp_fe_pose_update(self.h, len(meas), _p(meas), estimator, float(override_sigma), _p(out), _p(outlier)))
        return out[0], outlier[:len(meas)]

    def debug_scores(self, slot, level):
        w, h = self.cfg.width >> level, self.cfg.height >> level
        o = np.zeros((h, w), np.uint8)
        check(self.L.mcp_fe_debug_scores(self.h, slot, level, _p(o)))
        return o

    def timing(self):
        t = FeTiming()
        check(self.L.mcp_fe_get_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in t._fields_ if k != "pad_"}
