"""GPU parity tests of the bundle adjuster: CUDA path (through the C ABI) vs the CPU oracle.

The comparator is a RESTATEMENT of the reference (oracle/ba_oracle.c), not the reference binary
(itself pinned against the reference's own src/ChainBundle.cc in tests/test_oracle_vs_ref.py; g2o's LM loop [3P]).  Tolerances: residuals/Jacobians 1e-10 relative (same fp64
formulae, different rounding order), LM step and iterates 1e-6 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

from mcptam_b200 import synth

pytestmark = pytest.mark.gpu


def _mk(name, seed=0, **kw):
    return synth.make_ba_config(name, seed=seed, **kw)


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def capi():
    from mcptam_b200 import capi
    capi.lib()
    return capi


@pytest.mark.parametrize("cfg,seed", [("tiny", 0), ("tiny", 3), ("cfg1", 0)])
def test_eval_and_jacobian_parity(capi, cfg, seed):
    from oracle.oracle import OracleBA
    prob = _mk(cfg, seed)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    eg, cg = g.eval()
    eo, co = o.eval()
    assert np.allclose(eg, eo, rtol=1e-10, atol=1e-9)
    assert np.allclose(cg, co, rtol=1e-10, atol=1e-9)
    J = g.jacobians()
    rng = np.random.default_rng(seed)
    for m in rng.choice(prob.n_meas, min(prob.n_meas, 200), replace=False):
        jo, js, jp = o.jacobians(m)
        ref = np.concatenate([jo[0].ravel(), js[0].ravel(), jp.ravel()])
        scale = max(np.abs(ref).max(), 1.0)
        assert np.abs(J[m] - ref).max() <= 1e-10 * scale, (m, J[m], ref)


@pytest.mark.parametrize("cfg,seed,lam", [("tiny", 0, 10.0), ("tiny", 1, 1e-3), ("cfg1", 0, 50.0), ("cfg2", 0, 100.0)])
def test_lm_step_parity(capi, cfg, seed, lam):
    """One LM trial from identical (state, lambda, sigma^2): update vector within 1e-6 relative."""
    from oracle.oracle import OracleBA
    prob = _mk(cfg, seed)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    rc, d_o, sig_o, chi_o = o.lm_step(lam, -1.0, 0)
    assert rc == 0
    d_g, sig_g, chi_g = g.lm_step(lam, -1.0)
    assert abs(sig_g - sig_o) <= 1e-12 * sig_o            # exact upper median -> identical sigma
    assert abs(chi_g - chi_o) <= 1e-9 * abs(chi_o)
    nc = 6 * o.n_pose_var
    assert _rel(d_g[:nc], d_o[:nc]) < 1e-6
    assert _rel(d_g[nc:], d_o[nc:]) < 1e-6
    # and against the dense non-Schur solve of the full (poses+points) system on the small maps
    if cfg == "tiny":
        rc, d_f, _, _ = o.lm_step(lam, -1.0, 1)
        assert rc == 0 and _rel(d_g, d_f) < 1e-6


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("cfg,seed,lam", [("tiny", 0, 10.0), ("cfg1", 0, 50.0), ("cfg2", 0, 100.0)])
def test_schur_variants_parity(capi, monkeypatch, cfg, seed, lam, mode):
    """The non-default Schur reductions (MCP_BA_SCHUR=0 row-wise k_schur_rows, =2 staged pair gathers; the mode is
    read by mcp_ba_load) give the same LM update as the oracle and as the default TMA pair kernel."""
    from oracle.oracle import OracleBA
    prob = _mk(cfg, seed)
    o = OracleBA(prob)
    rc, d_o, _, chi_o = o.lm_step(lam, -1.0, 0)
    assert rc == 0
    monkeypatch.delenv("MCP_BA_SCHUR", raising=False)
    g1 = capi.BaHandle()
    g1.load(prob)
    d_1, _, _ = g1.lm_step(lam, -1.0)
    monkeypatch.setenv("MCP_BA_SCHUR", mode)
    g = capi.BaHandle()
    g.load(prob)
    d_g, _, chi_g = g.lm_step(lam, -1.0)
    assert abs(chi_g - chi_o) <= 1e-9 * abs(chi_o)
    assert _rel(d_g, d_o) < 1e-6
    assert _rel(d_g, d_1) < 1e-9


def _variant(prob, variant):
    """The other chain shapes the reference builds: fixed world points on a one-link [world] chain
    (src/BundleAdjusterMulti.cc:143-149, chi2 negated src/ChainBundle.cc:413-414) and BundleAdjusterSingle's one-link
    keyframe chains (src/BundleAdjusterSingle.cc:83-151)."""
    if "fixed" in variant:
        prob = synth.with_fixed_points(prob, 0.15)
    if "single" in variant:
        prob = synth.as_single_link(prob)
    return prob


@pytest.mark.parametrize("cfg,seed,iters", [("tiny", 0, 12), ("tiny", 2, 12), ("cfg1", 0, 10), ("cfg1", 1, 10), ("cfg2", 0, 10)])
def test_compute_parity(capi, monkeypatch, cfg, seed, iters):
    """Full Compute through the DEFAULT path (3 speculative candidates, fused multi-candidate Schur pass, look-ahead,
    PDL) vs the oracle -- including the benchmarked configuration (cfg2, 10 iterations = one bench.py step)."""
    from oracle.oracle import OracleBA
    for k in ("MCP_BA_SPECULATE", "MCP_BA_FUSE_SCHUR", "MCP_BA_LOOKAHEAD", "MCP_BA_PDL", "MCP_BA_SCHUR", "MCP_BA_GRAPH"):
        monkeypatch.delenv(k, raising=False)
    prob = _mk(cfg, seed)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    rc_o, st_o = o.compute(iters)
    rc_g, st_g = g.compute(iters)
    assert rc_g == rc_o
    assert st_g.iterations == st_o.iterations and st_g.total_trials == st_o.total_trials
    assert st_g.converged == st_o.converged
    assert _rel(g.poses(), o.poses()) < 1e-6
    assert _rel(g.points(), o.points()) < 1e-6
    assert abs(st_g.sigma_sq - st_o.sigma_sq) <= 1e-6 * st_o.sigma_sq
    assert abs(st_g.chi2_after - st_o.chi2_after) <= 1e-6 * st_o.chi2_after
    assert abs(st_g.chi2_before - st_o.chi2_before) <= 1e-9 * st_o.chi2_before
    assert abs(st_g.lambda_ - st_o.lambda_) <= 1e-5 * st_o.lambda_
    assert sorted(g.outliers().tolist()) == sorted(o.outliers().tolist())


@pytest.mark.parametrize("cfg,seed,iters", [("cfg1", 0, 8), ("cfg2", 0, 4)])
def test_speculation_variants_give_the_same_iterates(capi, monkeypatch, cfg, seed, iters):
    """Sequential trials (MCP_BA_SPECULATE=1), 2 and 3 concurrent candidates, and the fused multi-candidate Schur
    pass on/off (MCP_BA_FUSE_SCHUR) walk through the same lambda sequence and end in the same state."""
    prob = _mk(cfg, seed)
    res = {}
    for name, env in [("seq", {"MCP_BA_SPECULATE": "1"}), ("spec2", {"MCP_BA_SPECULATE": "2"}), ("spec3", {}),
                      ("spec3_unfused", {"MCP_BA_FUSE_SCHUR": "0"}), ("spec4", {"MCP_BA_SPECULATE": "4"})]:
        for k in ("MCP_BA_SPECULATE", "MCP_BA_FUSE_SCHUR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g = capi.BaHandle()
        g.load(prob)
        rc, st = g.compute(iters)
        res[name] = (rc, st.total_trials, st.lambda_, st.chi2_after, g.poses().copy(), g.points().copy())
    ref = res["seq"]
    for name, r in res.items():
        assert r[0] == ref[0] and r[1] == ref[1], name
        assert abs(r[2] - ref[2]) <= 1e-9 * ref[2], name
        assert abs(r[3] - ref[3]) <= 1e-9 * ref[3], name
        assert _rel(r[4], ref[4]) < 1e-9 and _rel(r[5], ref[5]) < 1e-9, name


def test_two_step_and_reset(capi):
    """Compute twice on the same handle (BundleAdjusterMulti two-step, src/BundleAdjusterMulti.cc:210-223)."""
    from oracle.oracle import OracleBA
    prob = _mk("cfg1", 5)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    for n in (4, 6):
        rc_o, st_o = o.compute(n)
        rc_g, st_g = g.compute(n)
        assert rc_g == rc_o and st_g.total_trials == st_o.total_trials
        assert _rel(g.points(), o.points()) < 1e-6
    g.reset_state()
    assert np.array_equal(g.points(), prob.pt_xyz) and np.array_equal(g.poses(), prob.pose_Rt)


def test_abort_and_errors(capi):
    prob = _mk("tiny", 0)
    g = capi.BaHandle()
    g.load(prob)
    ab = np.ones(1, np.uint8)
    rc, st = g.compute(10, abort=ab)
    assert rc == 0 and st.iterations == 0          # aborted before the first step -> 0 (src/ChainBundle.cc:1365-1366)
    # movable second chain link is the calibration BA: not on the hot path
    import copy
    bad = copy.copy(prob)
    bad.pose_fixed = prob.pose_fixed.copy()
    bad.pose_fixed[prob.n_mkf] = 0
    g2 = capi.BaHandle()
    with pytest.raises(capi.McpError) as ei:
        g2.load(bad)
    assert ei.value.code == -103


def test_convergence_noise_free(capi):
    """Noise-free map: LM converges to the truth (gauge fixed by the first MKF; multi-camera rig fixes scale)."""
    prob = synth.make_ba_problem(n_cam=2, n_mkf=6, n_pt=300, seed=7, outlier_frac=0.0, pix_sigma=0.0)
    g = capi.BaHandle()
    g.load(prob)
    rc, st = g.compute(60)
    assert rc > 0
    assert np.abs(g.poses() - prob.truth_pose_Rt).max() < 1e-5
    assert np.abs(g.points() - prob.truth_pt_xyz).max() < 1e-4


@pytest.mark.parametrize("robust", [True, False])
def test_point_depth_covariance(capi, robust):
    """GetMaxCov(): median (2,2) point covariance with < 3 movable poses (src/ChainBundle.cc:1401-1448).  The CUDA
    path uses the Schur form V^-1 + Y^T S^-1 Y, the oracle a dense factorisation of the full Hessian."""
    from oracle.oracle import OracleBA
    prob = synth.make_ba_problem(n_cam=2, n_mkf=3, n_pt=150, seed=3)
    g = capi.BaHandle(use_robust=robust, use_tukey=robust)
    g.load(prob)
    o = OracleBA(prob, use_robust=robust, use_tukey=robust)
    rc_g, st_g = g.compute(6)
    rc_o, st_o = o.compute(6)
    assert rc_g == rc_o and st_g.total_trials == st_o.total_trials
    assert st_o.max_cov > 0 and abs(st_g.max_cov - st_o.max_cov) <= 1e-6 * st_o.max_cov
    # >= 3 movable poses: not attempted, the reference's "failed" value
    g2 = capi.BaHandle()
    g2.load(synth.make_ba_config("tiny", seed=0))
    rc, st = g2.compute(2)
    assert st.max_cov == 0


@pytest.mark.parametrize("variant", ["fixed", "single", "fixed+single"])
@pytest.mark.parametrize("cfg,seed", [("tiny", 0), ("cfg1", 2)])
def test_fixed_points_and_one_link_chains(capi, cfg, seed, variant):
    """Residuals, chi2 sign, Jacobians, one LM step and a full Compute on maps with fixed world points and/or
    one-link chains, vs the oracle."""
    from oracle.oracle import OracleBA
    prob = _variant(_mk(cfg, seed), variant)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    eg, cg = g.eval()
    eo, co = o.eval()
    assert np.allclose(eg, eo, rtol=1e-10, atol=1e-9)
    assert np.allclose(cg, co, rtol=1e-10, atol=1e-9)
    if "fixed" in variant:
        assert (co < 0).sum() > 0 and np.array_equal(cg < 0, co < 0)      # the sign flip really is exercised
    J = g.jacobians()
    rng = np.random.default_rng(seed)
    for m in rng.choice(prob.n_meas, min(prob.n_meas, 300), replace=False):
        jo, js, jp = o.jacobians(m)
        ref = np.concatenate([jo[0].ravel(), js[0].ravel(), jp.ravel()])
        assert np.abs(J[m] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1.0), (m, J[m], ref)
    rc, d_o, sig_o, chi_o = o.lm_step(20.0, -1.0, 0)
    assert rc == 0
    d_g, sig_g, chi_g = g.lm_step(20.0, -1.0)
    assert abs(sig_g - sig_o) <= 1e-12 * sig_o
    assert abs(chi_g - chi_o) <= 1e-9 * abs(chi_o)
    assert _rel(d_g, d_o) < 1e-6
    rc_o, st_o = o.compute(8)
    rc_g, st_g = g.compute(8)
    assert rc_g == rc_o and st_g.total_trials == st_o.total_trials and st_g.converged == st_o.converged
    assert _rel(g.poses(), o.poses()) < 1e-6 and _rel(g.points(), o.points()) < 1e-6
    assert abs(st_g.sigma_sq - st_o.sigma_sq) <= 1e-6 * st_o.sigma_sq
    assert abs(st_g.chi2_after - st_o.chi2_after) <= 1e-6 * abs(st_o.chi2_after)
    assert sorted(g.outliers().tolist()) == sorted(o.outliers().tolist())
    if "fixed" in variant:                                               # fixed points do not move
        fx = prob.pt_fixed.astype(bool)
        assert np.array_equal(g.points()[fx], prob.pt_xyz[fx])


def test_cfg4_parity(capi):
    """The 8-camera 1000 KF / 100k-point map (BASELINE.json configs[3]) on one GPU: one LM step and a 2-iteration
    Compute through the default path vs the oracle (~25 s of CPU)."""
    from oracle.oracle import OracleBA
    prob = _mk("cfg4", 0)
    g = capi.BaHandle()
    g.load(prob)
    o = OracleBA(prob)
    rc, d_o, sig_o, chi_o = o.lm_step(100.0, -1.0, 0)
    assert rc == 0
    d_g, sig_g, chi_g = g.lm_step(100.0, -1.0)
    assert abs(sig_g - sig_o) <= 1e-12 * sig_o
    assert abs(chi_g - chi_o) <= 1e-9 * abs(chi_o)
    nc = 6 * o.n_pose_var
    assert _rel(d_g[:nc], d_o[:nc]) < 1e-6 and _rel(d_g[nc:], d_o[nc:]) < 1e-6
    rc_o, st_o = o.compute(2)
    rc_g, st_g = g.compute(2)
    assert rc_g == rc_o and st_g.total_trials == st_o.total_trials
    assert _rel(g.poses(), o.poses()) < 1e-6 and _rel(g.points(), o.points()) < 1e-6
    assert abs(st_g.sigma_sq - st_o.sigma_sq) <= 1e-6 * st_o.sigma_sq
    assert abs(st_g.lambda_ - st_o.lambda_) <= 1e-5 * st_o.lambda_
    assert abs(st_g.chi2_after - st_o.chi2_after) <= 1e-6 * st_o.chi2_after
    assert sorted(g.outliers().tolist()) == sorted(o.outliers().tolist())
