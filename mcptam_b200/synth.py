"""Seeded synthetic maps and frames for the two hot paths.

This module is data generation only (numpy): camera rigs with Taylor (Scaramuzza) parameters,
multi-keyframe trajectories, map points, measurements, and textured 640x480 frames.  It also
contains a numpy restatement of TaylorCamera::RefreshParams / FindInvPolyUsingRoots
(reference src/TaylorCamera.cc:84-198, 489-604), used to fill the plain-C camera struct that
crosses the C ABI (include/mcptam_b200.h: McpTaylorCam).
"""
from __future__ import annotations

import ctypes
import bisect
import hashlib
import os
import tempfile
import dataclasses
import math

import numpy as np

MAX_INV_DEGREE = 30  # include/mcptam/TaylorCamera.h:74


class TaylorCamStruct(ctypes.Structure):
    """Memory layout shared by McpTaylorCam (include/mcptam_b200.h) and OraTaylorCam (oracle/oracle.h)."""

    _fields_ = [
        ("poly", ctypes.c_double * 5),
        ("center", ctypes.c_double * 2),
        ("affine", ctypes.c_double * 4),
        ("image_size", ctypes.c_double * 2),
        ("min_theta", ctypes.c_double),
        ("theta_mean", ctypes.c_double),
        ("theta_std", ctypes.c_double),
        ("n_inv", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("inv_poly", ctypes.c_double * 32),
    ]


def _polyval_low_first(c, x):
    """TaylorCamera::PolyVal (src/TaylorCamera.cc:472-485): coefficient of x^0 first."""
    val = 0.0
    for i in range(len(c) - 1, 0, -1):
        val += c[i]
        val *= x
    return val + c[0]


def taylor_camera(params9, calib_size=(640, 480), full_size=(640, 480), image_size=(640, 480)) -> TaylorCamStruct:
    """numpy restatement of TaylorCamera::RefreshParams for a live (non-calibration) camera."""
    p = np.asarray(params9, dtype=np.float64)
    poly = np.array([p[0], 0.0, p[1], p[2], p[3]])
    calib = np.asarray(calib_size, float)
    full = np.asarray(full_size, float)
    img = np.asarray(image_size, float)
    scale = img / full
    fs_center = np.array([p[4] - (calib[0] - full[0]) / 2, p[5] - (calib[1] - full[1]) / 2])
    center = fs_center * scale
    corner = np.maximum(fs_center, full - fs_center - 1)
    largest_radius = math.sqrt(float(corner @ corner))
    max_rho = 1.0 * largest_radius
    min_theta = math.atan(_polyval_low_first(poly, max_rho) / max_rho)

    # FindInvPolyUsingRoots (src/TaylorCamera.cc:489-604)
    theta_start = -math.pi / 2 + 0.001
    theta_end = math.pi / 2 - 0.001
    step = 0.01
    n_theta = int(math.ceil((theta_end - theta_start) / step)) + 1
    thetas = np.empty(n_theta)
    thetas[0] = theta_start
    for i in range(1, n_theta):
        thetas[i] = thetas[i - 1] + step
    th_ok, rho_ok = [], []
    for th in thetas:
        coeffs_high_first = [poly[4], poly[3], poly[2], poly[1] - math.tan(th), poly[0]]
        roots = np.roots(coeffs_high_first)
        real = [r.real for r in roots if abs(r.imag) < 1e-12]
        real = [r for r in real if not (r < 0.0 or r > max_rho)]
        if len(real) == 1:
            th_ok.append(th)
            rho_ok.append(real[0])
    th_ok = np.array(th_ok)
    rho_ok = np.array(rho_ok)
    if th_ok.size < 3:
        raise ValueError("camera polynomial has no valid theta range")
    mean = float(th_ok.sum() / th_ok.size)
    shifted = th_ok - mean
    std = math.sqrt(float(shifted @ shifted) / shifted.size)
    x = (th_ok - mean) / std
    inv = None
    deg = 2
    while deg <= MAX_INV_DEGREE:
        V = np.vander(x, deg + 1, increasing=True)
        a, *_ = np.linalg.lstsq(V, rho_ok, rcond=None)  # SVD back-substitution (TooN SVD::backsub)
        err = np.abs(V @ a - rho_ok).max()
        if err <= 1e-4:
            inv = a
            break
        deg += 1
    if inv is None:
        raise ValueError("no inverse polynomial of degree <= %d fits to 1e-4" % MAX_INV_DEGREE)

    cam = TaylorCamStruct()
    cam.poly[:] = poly.tolist()
    cam.center[:] = center.tolist()
    cam.affine[:] = [scale[0] * p[6], scale[1] * p[7], scale[0] * p[8], scale[1] * 1.0]
    cam.image_size[:] = img.tolist()
    cam.min_theta = min_theta
    cam.theta_mean = mean
    cam.theta_std = std
    cam.n_inv = len(inv)
    for i, v in enumerate(inv):
        cam.inv_poly[i] = float(v)
    return cam


def cam_project_np(cam: TaylorCamStruct, v):
    """Vectorised TaylorCamera::Project (src/TaylorCamera.cc:202-287). v: (...,3). Returns px (...,2), invalid (...)."""
    v = np.asarray(v, float)
    norm = np.sqrt(v[..., 0] ** 2 + v[..., 1] ** 2)
    safe = np.where(norm == 0, 1.0, norm)
    theta = np.where(norm == 0, math.pi / 2, np.arctan(v[..., 2] / safe))
    inv = np.array(cam.inv_poly[: cam.n_inv])
    xs = (theta - cam.theta_mean) / cam.theta_std
    rho = np.zeros_like(xs)
    for i in range(len(inv) - 1, 0, -1):
        rho = (rho + inv[i]) * xs
    rho = rho + inv[0]
    rho = np.where(norm == 0, 0.0, rho)
    c = np.where(norm == 0, 0.0, v[..., 0] / safe)
    s = np.where(norm == 0, 0.0, v[..., 1] / safe)
    u, w = c * rho, s * rho
    A = cam.affine
    px = np.stack([A[0] * u + A[1] * w + cam.center[0], A[2] * u + A[3] * w + cam.center[1]], -1)
    invalid = theta < cam.min_theta
    invalid |= ~((px[..., 0] >= 0) & (px[..., 0] < cam.image_size[0]) & (px[..., 1] >= 0) & (px[..., 1] < cam.image_size[1]))
    return px, invalid


def cam_unproject_np(cam: TaylorCamStruct, px):
    """TaylorCamera::UnProject (src/TaylorCamera.cc:319-346), vectorised."""
    px = np.asarray(px, float)
    A = np.array(cam.affine).reshape(2, 2)
    Ai = np.linalg.inv(A)
    d = px - np.array(cam.center)
    uv = d @ Ai.T
    rho = np.sqrt((uv ** 2).sum(-1))
    poly = np.array(cam.poly)
    z = np.zeros_like(rho)
    for i in range(4, 0, -1):
        z = (z + poly[i]) * rho
    z = z + poly[0]
    ray = np.concatenate([uv, z[..., None]], -1)
    return ray / np.linalg.norm(ray, axis=-1, keepdims=True)


# ---------------------------------------------------------------------------------------------
# SE3 helpers (numpy, generator side only)
# ---------------------------------------------------------------------------------------------
def so3_exp(w):
    w = np.asarray(w, float)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * (K @ K)


def rt_mul(a, b):
    """(R,t) product: a*b."""
    return a[0] @ b[0], a[0] @ b[1] + a[1]


def rt_inv(a):
    return a[0].T, -a[0].T @ a[1]


def rt_pack(rt):
    return np.concatenate([rt[0].reshape(9), rt[1].reshape(3)])


@dataclasses.dataclass
class BaProblem:
    """Flat arrays of one bundle-adjustment problem — the fields the C ABI takes (mcp_ba_load)."""

    cams: list            # list[TaylorCamStruct]
    pose_Rt: np.ndarray   # (n_pose,12) row-major R then t; MKF base poses first, then cam-from-base
    pose_fixed: np.ndarray  # (n_pose,) u8
    pt_xyz: np.ndarray    # (n_pt,3) in source-camera frame
    pt_chain: np.ndarray  # (n_pt,2) i32 pose ids [mkf, cam-extrinsic]
    pt_fixed: np.ndarray  # (n_pt,) u8
    meas_xy: np.ndarray   # (n_meas,2)
    meas_chain: np.ndarray  # (n_meas,2) i32
    meas_pt: np.ndarray   # (n_meas,) i32
    meas_noise: np.ndarray  # (n_meas,) f64  = LevelScale^2  (BundleAdjusterMulti.cc:196)
    meas_cam: np.ndarray  # (n_meas,) i32 camera model index
    n_mkf: int = 0
    truth_pose_Rt: np.ndarray | None = None
    truth_pt_xyz: np.ndarray | None = None

    @property
    def n_pose(self):
        return len(self.pose_Rt)

    @property
    def n_pt(self):
        return len(self.pt_xyz)

    @property
    def n_meas(self):
        return len(self.meas_pt)


DEFAULT_TAYLOR = (250.0, -1.2e-3, 6.0e-7, -1.0e-9)  # a0, a2, a3, a4 : +z looking, ~165 deg FOV at 640x480


def make_rig(n_cam, rng, image_size=(640, 480)):
    """C cameras on a ring (radius 0.1 m), yaw 360/C apart, optical axis (+z) pointing outward."""
    cams, extr = [], []
    for c in range(n_cam):
        a0, a2, a3, a4 = DEFAULT_TAYLOR
        params = [a0 * (1 + 0.01 * rng.standard_normal()), a2, a3, a4,
                  image_size[0] / 2 + 2 * rng.standard_normal(), image_size[1] / 2 + 2 * rng.standard_normal(),
                  1.0 + 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal(), 1e-3 * rng.standard_normal()]
        cams.append(taylor_camera(params, image_size, image_size, image_size))
        yaw = 2 * math.pi * c / n_cam
        # base frame: x forward, y left, z up.  camera: z = optical axis, x right, y down.
        fwd = np.array([math.cos(yaw), math.sin(yaw), 0.0])
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        R_base_from_cam = np.stack([right, -up, fwd], 1)
        t_base_from_cam = 0.1 * fwd
        extr.append(rt_inv((R_base_from_cam, t_base_from_cam)))  # cam-from-base
    return cams, extr


def make_ba_problem(n_cam=1, n_mkf=20, n_pt=1000, seed=0, mean_track=8.0, outlier_frac=0.02,
                    pose_sigma_t=0.02, pose_sigma_r=math.radians(0.5), depth_sigma=0.05,
                    pix_sigma=0.5) -> BaProblem:
    rng = np.random.default_rng(seed)
    cams, extr = make_rig(n_cam, rng)
    # trajectory: loop of radius 5 m, +-0.5 m vertical sine; first MKF at identity (fixed)
    base_from_world = []
    for m in range(n_mkf):
        ang = 2 * math.pi * m / max(n_mkf, 1) * 0.9
        pos = np.array([5 * math.sin(ang), 5 * (1 - math.cos(ang)), 0.5 * math.sin(2 * ang)])
        Rwb = so3_exp([0, 0, ang]) @ so3_exp(0.05 * rng.standard_normal(3) * (m > 0))
        base_from_world.append(rt_inv((Rwb, pos)))
    # points: shell 3-12 m from the path
    centre = np.array([0.0, 5.0, 0.0])
    pts = []
    while len(pts) < n_pt:
        ang = rng.uniform(0, 2 * math.pi)
        path = np.array([5 * math.sin(ang), 5 * (1 - math.cos(ang)), 0.0])
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        pts.append(path + d * rng.uniform(3, 12))
    pts = np.array(pts)
    del centre

    # visibility: every point has a "home" MKF; only MKFs within +-win of it are considered (co-visibility is local
    # in real maps, and this keeps the table at O(P * win * C) instead of O(P * M * C))
    win = min(n_mkf - 1, 12)
    home = rng.integers(0, n_mkf, n_pt)
    Rs = np.array([[rt_mul(extr[c], base_from_world[m])[0] for c in range(n_cam)] for m in range(n_mkf)])   # (M,C,3,3)
    ts = np.array([[rt_mul(extr[c], base_from_world[m])[1] for c in range(n_cam)] for m in range(n_mkf)])   # (M,C,3)
    offs = np.arange(-win, win + 1)
    mk_idx = np.clip(home[:, None] + offs[None, :], 0, n_mkf - 1)          # (P, 2win+1)
    nW = len(offs)
    px_all = np.zeros((n_pt, nW, n_cam, 2), np.float32)
    ok_all = np.zeros((n_pt, nW, n_cam), bool)
    for j in range(nW):
        m = mk_idx[:, j]
        first = np.ones(n_pt, bool) if j == 0 else (m != mk_idx[:, j - 1])   # clipped duplicates count once
        for c in range(n_cam):
            pc = np.einsum("pij,pj->pi", Rs[m, c], pts) + ts[m, c]
            px, invalid = cam_project_np(cams[c], pc)
            margin = (px[:, 0] > 8) & (px[:, 0] < cams[c].image_size[0] - 8) & (px[:, 1] > 8) & (px[:, 1] < cams[c].image_size[1] - 8)
            off_axis = np.hypot(pc[:, 0], pc[:, 1]) > 1e-3 * np.abs(pc[:, 2])
            px_all[:, j, c] = px
            ok_all[:, j, c] = (~invalid) & margin & off_axis & first

    meas_xy, meas_chain, meas_pt, meas_noise, meas_cam = [], [], [], [], []
    pt_chain = np.zeros((n_pt, 2), np.int32)
    pt_rel_true = np.zeros((n_pt, 3))
    keep_pt = np.zeros(n_pt, bool)
    level_p = np.array([0.5, 0.25, 0.15, 0.1])
    level_cdf = level_p.cumsum()
    level_cdf /= level_cdf[-1]
    level_cdf = level_cdf.tolist()
    Ls = np.maximum(2, np.rint(rng.gamma(4.0, mean_track / 4.0, n_pt)).astype(int))
    u_all = rng.random((n_pt, 4))
    for p in range(n_pt):
        okp = ok_all[p]
        vis = np.argwhere(okp)
        if len(vis) < 2:
            continue
        js, cs = vis[int(u_all[p, 0] * len(vis))]
        ms = int(mk_idx[p, js])
        # observers: the MKFs of the window closest to the source
        jv = np.unique(vis[:, 0])
        order = jv[np.argsort(np.abs(mk_idx[p, jv] - ms), kind="stable")]
        chosen = order[: Ls[p]]
        obs = []
        for j in chosen:
            cc = np.flatnonzero(okp[j])
            if j == js:
                sel = [cs]
                if len(cc) > 1 and rng.random() < 0.15:
                    sel.append(int(rng.choice(cc[cc != cs])))
            else:
                sel = [int(cc[int(rng.random() * len(cc))])]
                if len(cc) > 1 and rng.random() < 0.15:
                    sel.append(int(rng.choice(cc[cc != sel[0]])))
            obs += [(int(j), int(c)) for c in sel]
        if len(obs) < 2:
            continue
        keep_pt[p] = True
        pt_chain[p] = (ms, n_mkf + cs)
        pt_rel_true[p] = Rs[ms, cs] @ pts[p] + ts[ms, cs]
        for j, c in obs:
            m = int(mk_idx[p, j])
            lvl = bisect.bisect_right(level_cdf, rng.random())   # == rng.choice(4, p=level_p): same draw, same stream
            z = px_all[p, j, c].astype(np.float64) + rng.standard_normal(2) * pix_sigma * (1 << lvl)
            if rng.random() < outlier_frac and not (m == ms and c == cs):
                z = px_all[p, j, c].astype(np.float64) + rng.uniform(-30, 30, 2)
            meas_xy.append(z)
            meas_chain.append((m, n_mkf + c))
            meas_pt.append(p)
            meas_noise.append(float((1 << lvl) ** 2))
            meas_cam.append(c)

    # compact point indices
    remap = -np.ones(n_pt, np.int64)
    remap[keep_pt] = np.arange(keep_pt.sum())
    meas_pt = remap[np.array(meas_pt, np.int64)].astype(np.int32)

    truth_pose = np.array([rt_pack(x) for x in base_from_world] + [rt_pack(e) for e in extr])
    # initial state: perturb movable MKF poses and point depths
    init_pose = truth_pose.copy()
    for m in range(1, n_mkf):
        dR = so3_exp(rng.standard_normal(3) * pose_sigma_r)
        dt = rng.standard_normal(3) * pose_sigma_t
        R, t = base_from_world[m]
        init_pose[m] = rt_pack((dR @ R, dR @ t + dt))
    pose_fixed = np.zeros(n_mkf + n_cam, np.uint8)
    pose_fixed[0] = 1
    pose_fixed[n_mkf:] = 1
    truth_rel = pt_rel_true[keep_pt]
    init_rel = truth_rel * np.exp(rng.standard_normal((len(truth_rel), 1)) * depth_sigma)

    return BaProblem(
        cams=cams, pose_Rt=np.ascontiguousarray(init_pose), pose_fixed=pose_fixed,
        pt_xyz=np.ascontiguousarray(init_rel), pt_chain=np.ascontiguousarray(pt_chain[keep_pt]),
        pt_fixed=np.zeros(int(keep_pt.sum()), np.uint8),
        meas_xy=np.ascontiguousarray(np.array(meas_xy)), meas_chain=np.array(meas_chain, np.int32),
        meas_pt=meas_pt, meas_noise=np.array(meas_noise), meas_cam=np.array(meas_cam, np.int32),
        n_mkf=n_mkf, truth_pose_Rt=truth_pose, truth_pt_xyz=truth_rel)


BA_CONFIGS = {
    # BASELINE.json configs[0], [1], [3]
    "cfg1": dict(n_cam=1, n_mkf=20, n_pt=1000),
    "cfg2": dict(n_cam=4, n_mkf=50, n_pt=10000),
    "cfg4": dict(n_cam=8, n_mkf=125, n_pt=100000),
    "tiny": dict(n_cam=2, n_mkf=5, n_pt=120),
}


_ARRAY_FIELDS = ("pose_Rt", "pose_fixed", "pt_xyz", "pt_chain", "pt_fixed", "meas_xy", "meas_chain", "meas_pt", "meas_noise",
                 "meas_cam", "truth_pose_Rt", "truth_pt_xyz")


def _cache_path(name, seed):
    """The big maps take tens of seconds of per-point Python to generate; the arrays are cached on local disk
    (keyed by this file's contents, so a generator change never serves stale data).  MCP_SYNTH_CACHE=0 disables."""
    root = os.environ.get("MCP_SYNTH_CACHE", os.path.join(tempfile.gettempdir(), "mcptam_b200_synth"))
    if root in ("", "0"):
        return None
    with open(__file__, "rb") as f:
        tag = hashlib.sha1(f.read()).hexdigest()[:12]
    return os.path.join(root, "%s_seed%d_%s.npz" % (name, seed, tag))


def make_ba_config(name, seed=0, **kw) -> BaProblem:
    args = dict(BA_CONFIGS[name])
    args.update(kw)
    path = _cache_path(name, seed) if not kw and args["n_pt"] >= 5000 else None
    if path and os.path.exists(path):
        try:
            z = np.load(path)
            cams = []
            for raw in z["cams_raw"]:
                cams.append(TaylorCamStruct.from_buffer_copy(raw.tobytes()))
            return BaProblem(cams=cams, n_mkf=int(z["n_mkf"]), **{k: np.ascontiguousarray(z[k]) for k in _ARRAY_FIELDS})
        except Exception:
            pass                                   # unreadable / partial file: regenerate
    prob = make_ba_problem(seed=seed, **args)
    if path:
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = "%s.%d.tmp.npz" % (path, os.getpid())
            raw = np.stack([np.frombuffer(bytes(c), np.uint8) for c in prob.cams])
            np.savez(tmp, cams_raw=raw, n_mkf=prob.n_mkf, **{k: getattr(prob, k) for k in _ARRAY_FIELDS})
            os.replace(tmp, path)
        except Exception:
            pass
    return prob


# ---------------------------------------------------------------------------------------------
# frames for the front end
# ---------------------------------------------------------------------------------------------
def make_frame(w=640, h=480, seed=0, n_shapes=400, shift=(0.0, 0.0)) -> np.ndarray:
    """Multi-octave value noise + random high-contrast rectangles/discs (≈2-4 k FAST corners at L0)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float64)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    xx = xx + shift[0]
    yy = yy + shift[1]
    for octave, amp in ((64, 50.0), (32, 30.0), (16, 18.0), (8, 10.0)):
        gh, gw = h // octave + 3, w // octave + 3
        g = rng.uniform(-1, 1, (gh, gw))
        fx, fy = xx / octave, yy / octave
        ix = np.clip(np.floor(fx).astype(int), 0, gw - 2)
        iy = np.clip(np.floor(fy).astype(int), 0, gh - 2)
        ax, ay = fx - ix, fy - iy
        v = (g[iy, ix] * (1 - ax) + g[iy, ix + 1] * ax) * (1 - ay) + (g[iy + 1, ix] * (1 - ax) + g[iy + 1, ix + 1] * ax) * ay
        img += amp * v
    img += 128
    for _ in range(n_shapes):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        val = rng.choice([20.0, 60.0, 190.0, 235.0])
        if rng.random() < 0.6:
            hw, hh = rng.uniform(3, 18), rng.uniform(3, 18)
            m = (np.abs(xx - cx) < hw) & (np.abs(yy - cy) < hh)
        else:
            r = rng.uniform(3, 14)
            m = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
        img[m] = val
    img += rng.standard_normal((h, w)) * 1.5
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# Two-view scene for the epipolar point creation tests: a textured plane rendered through the Taylor camera
# ---------------------------------------------------------------------------------------------
def render_plane_view(cam: TaylorCamStruct, cam_from_world, texture, plane_z=4.0, texel=0.0125):
    """Image of the world plane z = plane_z (texture centred on the world z axis, `texel` metres per texture pixel) seen
    by `cam` at pose cam_from_world (12 doubles: row-major R, t).  Bilinear texture lookup, 0 outside the texture."""
    w, h = int(cam.image_size[0]), int(cam.image_size[1])
    rt = np.asarray(cam_from_world, np.float64).reshape(-1)
    R, t = rt[:9].reshape(3, 3), rt[9:]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    rays_c = cam_unproject_np(cam, np.stack([xx, yy], -1))
    rays_w = rays_c @ R                                   # R^T applied to every ray
    origin = -R.T @ t
    with np.errstate(divide="ignore", invalid="ignore"):
        lam = (plane_z - origin[2]) / rays_w[..., 2]
    pw = origin + lam[..., None] * rays_w
    th, tw = texture.shape
    u = pw[..., 0] / texel + tw / 2.0
    v = pw[..., 1] / texel + th / 2.0
    ok = (lam > 0) & (u >= 0) & (u < tw - 1) & (v >= 0) & (v < th - 1)
    u = np.where(ok, u, 0.0); v = np.where(ok, v, 0.0)
    iu, iv = np.floor(u).astype(int), np.floor(v).astype(int)
    fu, fv = u - iu, v - iv
    tex = texture.astype(np.float64)
    val = (tex[iv, iu] * (1 - fu) * (1 - fv) + tex[iv, iu + 1] * fu * (1 - fv) + tex[iv + 1, iu] * (1 - fu) * fv + tex[iv + 1, iu + 1] * fu * fv)
    return np.where(ok, np.clip(np.rint(val), 0, 255), 0).astype(np.uint8)


def make_stereo_scene(seed=0, baseline=0.5, plane_z=4.0, yaw=-0.04):
    """Two cameras looking at a textured plane: returns dict(cam_a, cam_b, cfw_a, cfw_b (12 doubles), img_a, img_b,
    plane_z).  Camera A sits at the world origin looking along +z."""
    rng = np.random.default_rng(seed)
    cams, _ = make_rig(2, rng)
    texture = make_frame(w=1280, h=960, seed=seed + 11, n_shapes=500)
    cfw_a = np.concatenate([np.eye(3).reshape(-1), np.zeros(3)])
    Rb = so3_exp(np.array([0.01, yaw, 0.015]))
    centre_b = np.array([baseline, 0.03, -0.05])
    cfw_b = np.concatenate([Rb.reshape(-1), -Rb @ centre_b])
    return dict(cam_a=cams[0], cam_b=cams[1], cfw_a=cfw_a, cfw_b=cfw_b, plane_z=plane_z,
                img_a=render_plane_view(cams[0], cfw_a, texture, plane_z), img_b=render_plane_view(cams[1], cfw_b, texture, plane_z))


# ---------------------------------------------------------------------------------------------
# Variants of a problem that exercise the other chain shapes the reference builds
# ---------------------------------------------------------------------------------------------
def _rt_unpack(row):
    row = np.asarray(row, np.float64)
    return row[:9].reshape(3, 3), row[9:]


def with_fixed_points(prob: BaProblem, frac=0.1, seed=0) -> BaProblem:
    """Turns a fraction of the points into FIXED world points the way BundleAdjusterMulti marshals `point.mbFixed`
    (src/BundleAdjusterMulti.cc:143-149): one extra fixed identity "world" pose, a ONE-link chain [world] and the
    world position as coordinates.  Their chi2 enters the robust statistics negated (src/ChainBundle.cc:413-414).
    The positions come from the ground truth (a fixed point is one the map trusts)."""
    rng = np.random.default_rng(seed + 7919)
    n_pt = prob.n_pt
    pick = np.flatnonzero(rng.random(n_pt) < frac)
    world_id = prob.n_pose
    pose_Rt = np.vstack([prob.pose_Rt, np.concatenate([np.eye(3).reshape(-1), np.zeros(3)])[None]])
    truth_pose = np.vstack([prob.truth_pose_Rt, pose_Rt[-1:]]) if prob.truth_pose_Rt is not None else None
    pose_fixed = np.concatenate([prob.pose_fixed, np.ones(1, np.uint8)])
    pt_xyz, pt_chain, pt_fixed = prob.pt_xyz.copy(), prob.pt_chain.copy(), prob.pt_fixed.copy()
    truth_pt = prob.truth_pt_xyz.copy() if prob.truth_pt_xyz is not None else None
    src_pose = prob.truth_pose_Rt if prob.truth_pose_Rt is not None else prob.pose_Rt
    src_rel = prob.truth_pt_xyz if prob.truth_pt_xyz is not None else prob.pt_xyz
    for p in pick:
        m, c = prob.pt_chain[p]
        T = rt_mul(_rt_unpack(src_pose[c]), _rt_unpack(src_pose[m])) if c >= 0 else _rt_unpack(src_pose[m])
        Ri, ti = rt_inv(T)
        w = Ri @ src_rel[p] + ti
        pt_xyz[p] = w
        if truth_pt is not None:
            truth_pt[p] = w
        pt_chain[p] = (world_id, -1)
        pt_fixed[p] = 1
    return dataclasses.replace(prob, pose_Rt=np.ascontiguousarray(pose_Rt), pose_fixed=pose_fixed, pt_xyz=pt_xyz,
                               pt_chain=pt_chain, pt_fixed=pt_fixed, truth_pose_Rt=truth_pose, truth_pt_xyz=truth_pt)


def as_single_link(prob: BaProblem) -> BaProblem:
    """The same map marshalled the way BundleAdjusterSingle does (src/BundleAdjusterSingle.cc:83-151): every keyframe
    carries its own CamFromWorld pose, measurement and point chains have ONE link.  (Each camera of a multi-keyframe
    becomes an independent pose, so the optimum differs from the rig-constrained problem; it is the chain shape that
    matters here.)  Fixed world points (one-link chains onto the world pose) are kept as they are."""
    n_mkf = prob.n_mkf
    key = {}
    poses, truth, fixed = [], [], []

    def kf_id(m, c):
        if c < 0:
            k = (int(m), -1)
        else:
            k = (int(m), int(c))
        if k not in key:
            key[k] = len(poses)
            if c < 0:
                poses.append(prob.pose_Rt[m]); fixed.append(prob.pose_fixed[m])
                truth.append(prob.truth_pose_Rt[m] if prob.truth_pose_Rt is not None else prob.pose_Rt[m])
            else:
                poses.append(rt_pack(rt_mul(_rt_unpack(prob.pose_Rt[c]), _rt_unpack(prob.pose_Rt[m]))))
                fixed.append(prob.pose_fixed[m])
                tp = prob.truth_pose_Rt if prob.truth_pose_Rt is not None else prob.pose_Rt
                truth.append(rt_pack(rt_mul(_rt_unpack(tp[c]), _rt_unpack(tp[m]))))
        return key[k]

    # keyframes in (mkf, cam) order so that the fixed first MKF's keyframes come first
    pairs = sorted({(int(a), int(b)) for a, b in prob.meas_chain} | {(int(a), int(b)) for a, b in prob.pt_chain})
    for m, c in pairs:
        kf_id(m, c)
    meas_chain = np.array([(kf_id(a, b), -1) for a, b in prob.meas_chain], np.int32)
    pt_chain = np.array([(kf_id(a, b), -1) for a, b in prob.pt_chain], np.int32)
    del n_mkf
    return dataclasses.replace(prob, pose_Rt=np.ascontiguousarray(np.array(poses)), pose_fixed=np.array(fixed, np.uint8),
                               meas_chain=meas_chain, pt_chain=pt_chain, n_mkf=len(poses),
                               truth_pose_Rt=np.ascontiguousarray(np.array(truth)))
