"""Builds the C++ host mirror (libmcptam_host.so) and its test executable against libmcptam_b200.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "_build")
SRCS = ["TaylorCamera.cc", "ChainBundle.cc", "BundleAdjusterCuda.cc", "FrontEnd.cc", "MapIO.cc", "Epipolar.cc"]


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libmcptam_host.so")
    exe = os.path.join(OUT, "test_host")
    core = os.path.join(OUT, "libmcptam_b200.so")
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cc", ".h"))] + [os.path.join(HERE, "shim", f) for f in os.listdir(os.path.join(HERE, "shim"))]
    stale = force or not (os.path.exists(lib) and os.path.exists(exe)) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps)
    if not stale:
        return exe
    cxx = ["g++", "-O2", "-std=c++14", "-fPIC", "-Wall", "-Wno-unused-function"]
    subprocess.check_call(cxx + ["-shared", "-o", lib] + [os.path.join(HERE, s) for s in SRCS] + ["-L" + OUT, "-lmcptam_b200", "-Wl,-rpath,$ORIGIN"])
    subprocess.check_call(cxx + ["-o", exe, os.path.join(HERE, "test_host.cc"), "-L" + OUT, "-lmcptam_host", "-lmcptam_b200", "-Wl,-rpath,$ORIGIN",
                                 "-Wl,-rpath-link," + OUT])
    return exe


if __name__ == "__main__":
    print(build(force=True))
