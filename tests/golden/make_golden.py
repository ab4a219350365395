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


def epipolar_candidates(sc, per_level=8):
    """The candidates the epipolar fixture / tests use: FAST corners of the source view, seeded choice per level."""
    pa = ora.pyramid(sc["img_a"])
    rng = np.random.default_rng(4)
    cand = []
    for level in (0, 1, 2):
        cor = ora.level_corners(pa[level])["corners"]
        h, w = pa[level].shape
        cor = cor[(cor[:, 0] > 20) & (cor[:, 0] < w - 20) & (cor[:, 1] > 20) & (cor[:, 1] < h - 20)]
        for i in rng.choice(len(cor), per_level, replace=False):
            cand.append((level, int(cor[i][0]), int(cor[i][1])))
    return cand


def epipolar_fixture():
    from oracle import epipolar as epi
    sc = synth.make_stereo_scene(0)
    pa, pb = ora.pyramid(sc["img_a"]), ora.pyramid(sc["img_b"])
    lb = [ora.level_corners(x) for x in pb]
    cand = epipolar_candidates(sc)
    rows = []
    for level, x, y in cand:
        r = epi.add_point_epipolar(sc["cam_a"], sc["cam_b"], sc["cfw_a"], sc["cfw_b"], pa, pb, lb, level, (x, y))
        rows.append((int(r["ok"]), r.get("n_steps", -1), r.get("n_matches", -1), r.get("best", -1), r.get("best_score", -1)) +
                    tuple(r["world"] if r["ok"] else (0.0, 0.0, 0.0)) + tuple(r["subpix"] if r["ok"] else (0.0, 0.0)))
    np.savez_compressed(os.path.join(HERE, "epipolar_seed0.npz"), cand=np.array(cand, np.int32), rows=np.array(rows, np.float64),
                        img_a_sum=int(sc["img_a"].astype(np.int64).sum()), img_b_sum=int(sc["img_b"].astype(np.int64).sum()))


if __name__ == "__main__":
    ba_fixture()
    fe_fixture()
    epipolar_fixture()
    print("golden fixtures written to", HERE)
