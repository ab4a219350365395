"""On-disk map dump (MapMakerBase::DumpToFile, src/MapMakerBase.cc:475-579) written / read by the C++ host mirror
(mcptam_b200/host/MapIO.*).  CPU only: the C++ self test checks the round trip, this file parses the dump
independently and checks the record structure the reference writes."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mcptam_b200", "host")


def _build(tmp_path):
    exe = str(tmp_path / "test_mapio")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-o", exe, os.path.join(HOST, "MapIO.cc"), os.path.join(HOST, "test_mapio.cc")])
    return exe


def _records(path):
    lines = open(path).read().split("\n")
    assert lines[-1] == "% The end"                       # no trailing newline, as the reference leaves it
    return lines, [l for l in lines if l and not l.startswith("%")]


def test_dump_round_trip_and_format(tmp_path):
    exe = _build(tmp_path)
    dump = str(tmp_path / "map.txt")
    out = subprocess.run([exe, dump], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "MAPIO_TEST OK" in out.stdout, out.stdout[-2000:]
    lines, rec = _records(dump)
    # the comment block of every section is where the reference writes it
    assert lines[0] == "% Camera poses in MKF frame, format:" and lines[1] == "% Total number of cameras"
    assert sum(l.startswith("%") for l in lines) == 13
    it = iter(rec)
    n_cam = int(next(it))
    cams = [next(it).split(", ") for _ in range(n_cam)]
    n_mkf = int(next(it))
    mkfs = [next(it).split(", ") for _ in range(n_mkf)]
    n_pt = int(next(it))
    pts = [next(it).split(", ") for _ in range(n_pt)]
    n_meas = int(next(it))
    meas = [next(it).split(", ") for _ in range(n_meas)]
    assert list(it) == []
    names = [c[0] for c in cams]
    assert names == sorted(names) and all(len(c) == 8 for c in cams)       # std::map order of the first MKF's keyframes
    for r in cams + mkfs:
        q = np.array(r[4:8], float)
        assert abs(np.linalg.norm(q) - 1) < 1e-5                            # 6 significant digits
    assert [int(m[0]) for m in mkfs] == list(range(n_mkf))
    assert [int(p[0]) for p in pts] == list(range(n_pt))
    assert all(0 <= int(p[4]) < n_mkf and p[5] in names and len(p) == 6 for p in pts)
    assert all(len(m) == 6 and 0 <= int(m[0]) < n_mkf and m[1] in names and 0 <= int(m[2]) < n_pt for m in meas)
    assert set(int(m[5]) for m in meas) <= {1, 4, 16, 64}                   # LevelScale(level)^2
    # measurements are grouped by MKF in list order, inside an MKF by camera name
    keys = [(int(m[0]), m[1]) for m in meas]
    assert keys == sorted(keys)
    # a keyframe measures a point at most once
    assert len(set((m[0], m[1], m[2]) for m in meas)) == n_meas
    # the second dump (of the re-loaded map) has the same records up to the last printed digit
    _, rec2 = _records(dump + ".2")
    assert len(rec2) == len(rec)
    for a, b in zip(rec, rec2):
        fa, fb = a.split(", "), b.split(", ")
        assert len(fa) == len(fb)
        for x, y in zip(fa, fb):
            try:
                assert abs(float(x) - float(y)) <= 2e-5 * max(1.0, abs(float(x)))
            except ValueError:
                assert x == y
