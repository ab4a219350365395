"""Point-sharded multi-GPU BA == single-GPU BA (needs >= 2 CUDA devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_matches_single():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "ba_multi.py"), "cfg1", "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MULTI")]
    assert line and line[0].endswith("OK"), out.stdout[-2000:]
    ab = [l for l in out.stdout.splitlines() if l.startswith("MULTI_ABORT")]
    assert ab and ab[0].endswith("OK"), out.stdout[-2000:]
