"""The C++ host mirror (reference-signature classes over the C ABI) end to end on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror():
    import importlib.util
    from mcptam_b200 import build as b
    b.build()
    spec = importlib.util.spec_from_file_location("build_host", os.path.join(ROOT, "mcptam_b200", "host", "build_host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    exe = mod.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "HOST_TEST OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
