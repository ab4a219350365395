"""CPU checks of the C++ host mirror (mcptam_b200/host): the TaylorCamera class (restating src/TaylorCamera.cc:84-383,
including the inverse-polynomial fit :489-604) against the independent Python restatement in mcptam_b200/synth.py and the
C oracle's camera functions.  Two restatements of the same reference code written separately must agree."""
import ctypes as C
import importlib.util
import math
import os

import numpy as np
import pytest

from mcptam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    from mcptam_b200 import capi
    capi.lib()
    spec = importlib.util.spec_from_file_location("build_host", os.path.join(ROOT, "mcptam_b200", "host", "build_host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return C.CDLL(os.path.join(ROOT, "mcptam_b200", "_build", "libmcptam_host.so"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


PARAMS = [
    [250.0, -1.2e-3, 6.0e-7, -1.0e-9, 320.0, 240.0, 1.0, 0.0, 0.0],              # synth.DEFAULT_TAYLOR
    [252.3, -1.25e-3, 6.4e-7, -0.9e-9, 322.5, 238.1, 1.001, 0.0007, -0.0011],
    [230.0, -1.0e-3, 5.0e-7, -0.8e-9, 316.0, 243.0, 0.998, -0.002, 0.001],
]


@pytest.mark.parametrize("params", PARAMS)
def test_camera_mirror_matches_the_python_restatement(host, params):
    from oracle import oracle as O
    ref = synth.taylor_camera(params, (640, 480), (640, 480), (640, 480))
    got = synth.TaylorCamStruct()
    p9 = np.array(params, np.float64)
    assert host.mcp_host_camera_abi(_p(p9), 640, 480, C.byref(got)) == 0
    assert np.allclose(got.poly[:], ref.poly[:], rtol=0, atol=0)
    assert np.allclose(got.center[:], ref.center[:], rtol=1e-15) and np.allclose(got.affine[:], ref.affine[:], rtol=1e-15, atol=1e-18)
    assert got.image_size[:] == ref.image_size[:]
    assert abs(got.min_theta - ref.min_theta) < 1e-14
    assert abs(got.theta_mean - ref.theta_mean) < 1e-12 and abs(got.theta_std - ref.theta_std) < 1e-12
    assert got.n_inv == ref.n_inv                                   # same degree passes the 1e-4 pixel fit test
    # the fitted coefficients come out of different least-squares solvers (TooN-style SVD vs LAPACK): compare the
    # polynomials where they are used, in pixels
    xs = np.linspace(-1.7, 1.7, 400)
    pg = sum(got.inv_poly[i] * xs ** i for i in range(got.n_inv))
    pr = sum(ref.inv_poly[i] * xs ** i for i in range(ref.n_inv))
    assert np.abs(pg - pr).max() < 1e-6
    # Project / derivatives / UnProject of the C++ class vs the C oracle evaluated on the PYTHON restatement's record
    rng = np.random.default_rng(0)
    seen = synth.cam_unproject_np(ref, rng.uniform([2, 2], [638, 478], (250, 2))) * rng.uniform(0.5, 20.0, (250, 1))     # in view
    pts = np.concatenate([seen, -seen[:30], rng.standard_normal((40, 3)), [[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]]])            # + behind / anywhere / on the axis
    n = len(pts)
    px, dv, inv, opa = np.zeros((n, 2)), np.zeros((n, 4)), np.zeros(n, np.int32), C.c_double()
    assert host.mcp_host_camera_project(_p(p9), 640, 480, n, _p(np.ascontiguousarray(pts)), _p(px), _p(dv), _p(inv), C.byref(opa)) == 0
    n_valid = 0
    for i in range(n):
        rpx, rd = np.zeros(2), np.zeros(4)
        rinv = O.lib().ora_cam_project(C.byref(ref), O._p(np.ascontiguousarray(pts[i])), O._p(rpx), O._p(rd))
        assert bool(rinv) == bool(inv[i]), (i, pts[i])
        if not rinv:
            n_valid += 1
            assert np.allclose(px[i], rpx, rtol=0, atol=2e-6)               # two fits of the same 1e-4-px-accurate polynomial
            assert np.allclose(dv[i], rd, rtol=1e-6, atol=1e-6)
    assert n_valid > 100
    from oracle import epipolar as E
    assert abs(opa.value - E.one_pixel_angle(ref)) < 1e-12
    pix = rng.uniform([5, 5], [635, 475], (200, 2))
    rays = np.zeros((200, 3))
    assert host.mcp_host_camera_unproject(_p(p9), 640, 480, 200, _p(np.ascontiguousarray(pix)), _p(rays)) == 0
    assert np.allclose(rays, synth.cam_unproject_np(ref, pix), rtol=0, atol=1e-14)
    # and the round trip through the class itself stays within the fit tolerance
    back, dv2, inv2 = np.zeros((200, 2)), np.zeros((200, 4)), np.zeros(200, np.int32)
    host.mcp_host_camera_project(_p(p9), 640, 480, 200, _p(np.ascontiguousarray(rays)), _p(back), _p(dv2), _p(inv2), None)
    ok = inv2 == 0
    assert ok.sum() > 150 and np.abs(back[ok] - pix[ok]).max() < 2e-4


def test_bad_camera_is_rejected(host):
    out = synth.TaylorCamStruct()
    for bad in ([1.0, -1.2e-3, 6e-7, -1e-9, 320.0, 240.0, 1.0, 0.0, 0.0],         # no inverse polynomial fits to 1e-4 px
                [250.0, -1.2e-3, 6e-7, 1e-5, 320.0, 240.0, 1.0, 0.0, 0.0], [250.0, -0.05, 0.0, 0.0, 320.0, 240.0, 1.0, 0.0, 0.0]):
        b = np.array(bad)
        assert host.mcp_host_camera_abi(_p(b), 640, 480, C.byref(out)) != 0
        with pytest.raises(ValueError):
            synth.taylor_camera(bad, (640, 480), (640, 480), (640, 480))


def test_bundle_adjuster_marshalling_equals_the_reference_loop(host, tmp_path):
    """BundleAdjusterCuda::Marshal (parallel per-keyframe gather, hash-map lookups, pooled arrays) hands mcp_ba_load exactly the arrays of
    the loop as src/BundleAdjusterMulti.cc:83-203 writes it (host/test_marshal_cpu.cc; no device call involved)."""
    import subprocess
    build = os.path.join(ROOT, "mcptam_b200", "_build")
    exe = str(tmp_path / "test_marshal")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-Wno-unused-function", "-o", exe,
                           os.path.join(ROOT, "mcptam_b200", "host", "test_marshal_cpu.cc"), "-L" + build, "-lmcptam_host", "-lmcptam_b200",
                           "-Wl,-rpath," + build, "-lpthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MARSHAL_TEST OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
