"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/mcptam_b200.h
declares, and the product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mcptam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mcptam_b200 import capi
    L = capi.lib()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.mcp_abi_version() >= 1


def test_struct_layouts_match_header():
    from mcptam_b200 import capi, synth
    assert C.sizeof(synth.TaylorCamStruct) == 5 * 8 + 2 * 8 + 4 * 8 + 2 * 8 + 3 * 8 + 8 + 32 * 8
    assert C.sizeof(capi.PatchReq) == 72 and C.sizeof(capi.PatchRes) == 48
    assert C.sizeof(capi.BaStats) == 6 * 4 + 7 * 8 + 8


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from mcptam_b200 import capi
    with pytest.raises(capi.McpError) as ei:
        capi.BaHandle()
    assert ei.value.code == -104 and "no CPU fallback" in str(ei.value)
    with pytest.raises(capi.McpError) as ei:
        capi.FeHandle()
    assert ei.value.code == -104


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mcptam_b200/ may import, link or load it."""
    for d, _, files in os.walk(os.path.join(ROOT, "mcptam_b200")):
        if "_build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(d, f)


def test_partition_balanced_and_complete():
    from mcptam_b200 import capi, synth
    prob = synth.make_ba_config("cfg1", 0)
    for world in (1, 2, 4, 8):
        part = capi.ba_partition(prob.n_pt, prob.meas_pt, world)
        assert part[0] == 0 and part[-1] == prob.n_pt and (np.diff(part) >= 0).all()
        counts = np.bincount(prob.meas_pt, minlength=prob.n_pt)
        loads = [counts[part[r]:part[r + 1]].sum() for r in range(world)]
        assert max(loads) <= 1.25 * (sum(loads) / world) + 50
