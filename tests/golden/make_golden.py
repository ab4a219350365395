"""Generates the small golden fixtures under tests/golden/ from the CPU oracle (oracle/*.c).

The reference (aharmat/mcptam) ships no golden vectors and cannot be built or imported here, so these
fixtures pin the ORACLE (regression guard) and give the GPU tests fixed inputs/outputs that travel to the GPU
box.  Parity with the reference binary itself stays unpinned (see oracle/oracle.h).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mcptam_b200 import synth  # noqa: E402
from oracle import oracle as ora  # noqa: E402
from oracle.oracle import OracleBA  # noqa: E402


def ba_fixture():
    prob = synth.make_ba_config("tiny", seed=0)
    o = OracleBA(prob)
    e, c = o.eval()
    J = np.stack([np.concatenate([x.ravel() for x in (o.jacobians(m)[0][0], o.jacobians(m)[1][0], o.jacobians(m)[2])]) for m in range(prob.n_meas)])
    rc, delta, sig, chi = o.lm_step(10.0, -1.0, 0)
    rc, st = o.compute(8)
    np.savez_compressed(os.path.join(HERE, "ba_tiny_seed0.npz"), err=e, chi2=c, jac=J, lm_delta=delta, lm_sigma_sq=sig, lm_chi2=chi,
                        poses=o.poses(), points=o.points(), iterations=st.iterations, total_trials=st.total_trials,
                        sigma_sq=st.sigma_sq, chi2_after=st.chi2_after, outliers=o.outliers())


def fe_fixture():
    img = synth.make_frame(w=320, h=240, seed=5, n_shapes=120)
    pyr = ora.pyramid(img)
    out = {"img": img}
    for l, im in enumerate(pyr):
        r = ora.level_corners(im)
        out["corners%d" % l] = r["corners"]
        out["lut%d" % l] = r["row_lut"]
        out["thresh%d" % l] = r["fast_thresh"]
        out["freq%d" % l] = r["fast_freq"]
    np.savez_compressed(os.path.join(HERE, "fe_320x240_seed5.npz"), **out)


if __name__ == "__main__":
    ba_fixture()
    fe_fixture()
    print("golden fixtures written to", HERE)
