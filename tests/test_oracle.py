"""CPU suite, part 1: the oracle against the golden vectors and the reference's structural known answers."""
import os

import numpy as np
import pytest

import c_oracle as co
import gync_oracle as o

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def rhs_gold():
    return np.load(os.path.join(G, "rhs_golden.npz"))


@pytest.fixture(scope="module")
def sol_gold():
    return np.load(os.path.join(G, "solution_golden.npz"))


def relerr(a, b):
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))


def test_rhs_python_oracle_vs_reference_text(rhs_gold):
    # golden F was produced by executing the reference's own source text (make_rhs_golden.py)
    assert relerr(o.gyncycle_rhs(rhs_gold["Y"], rhs_gold["P"]), rhs_gold["F"]) < 1e-11


def test_rhs_c_oracle_vs_reference_text(rhs_gold):
    F = np.array([co.rhs(y, p) for y, p in zip(rhs_gold["Y"], rhs_gold["P"])])
    assert relerr(F, rhs_gold["F"]) < 1e-12


def test_rhs_anchor_values():
    # SURVEY §8c sanity anchor at refy0 / refallparms
    f = o.gyncycle_rhs(o.REFY0, o.REFALLPARMS)
    assert np.allclose(f[[0, 5, 1, 6]], [1584.6267, 5292.1452, 0.030491, 0.600699], rtol=1e-5)


def test_structural_known_answers():
    # (i) the 21 Hill-exponent slots of p are never read by the RHS (model.jl:3)
    p = o.REFALLPARMS.copy()
    f0 = o.gyncycle_rhs(o.REFY0, p)
    p[o.HILLINDS] *= 7.3
    assert np.array_equal(f0, o.gyncycle_rhs(o.REFY0, p))
    # (ii) 82 sampled parameters, sample width 116 (test/gync.jl:14)
    assert o.SAMPLEDINDS.size == 82 and o.init_x().size == 116
    assert np.array_equal(o.allparms(o.REFPARMS), o.REFALLPARMS)
    # (v) llh_measerror: all-NaN -> -Inf, NaN-free zeros -> 0 (model.jl:158-163)
    assert o.llh_measerror(np.full((31, 4), np.nan), 0.1) == -np.inf
    assert o.llh_measerror(np.zeros((31, 4)), 0.1) == 0.0
    # Jacobian sparsity: 117 structural non-zeros
    assert (o.analytic_jac(o.REFY0, o.REFALLPARMS) != 0).sum() == 117


def test_negative_base_of_real_power_is_nan():
    y = o.REFY0.copy()
    y[3] = -1.0      # LH_R < 0: Julia raises DomainError -> NaN trajectory (gyncycle.jl:18-19,144)
    assert np.isnan(o.gyncycle_rhs(y, o.REFALLPARMS)).any()
    assert np.isnan(co.rhs(y, o.REFALLPARMS)).any()


def test_meastimes_and_duplicates():
    t = o.meastimes(31, 2, 28.0)
    assert t.size == 62 and t[0] == 1 and t[31] == 29 and np.sum(np.diff(np.sort(t)) == 0) == 3   # days 29..31 twice


def test_c_oracle_converged_reproduces_golden(sol_gold):
    idx = [0, 1, 5, 34]
    LL, tr = co.llh_batch(sol_gold["X"][idx], [sol_gold["data"][:, :, m] for m in range(4)], [0.1] * 4, 1, 1e-12, 1e-12,
                          want_traj=True)
    assert relerr(LL, sol_gold["LL"][idx]) < 1e-10
    assert relerr(tr, sol_gold["traj"][idx]) < 1e-9


def test_c_oracle_failure_semantics(sol_gold):
    # wide-set samples with negative initial values: NaN trajectory -> -Inf (model.jl:138-141)
    bad = np.flatnonzero(~np.isfinite(sol_gold["LL"][:, 0]))
    assert bad.size >= 1
    LL, tr = co.llh_batch(sol_gold["X"][bad[:2]], [sol_gold["data"][:, :, 0]], [0.1], 1, 1e-9, 1e-9, want_traj=True)
    assert np.all(LL == -np.inf) and np.isnan(tr).all()
    # all-missing patient table -> -Inf (model.jl:161)
    LL = co.llh_batch(sol_gold["X"][:1], [np.full((31, 4), np.nan)], [0.1], 0, 1e-4, 1e-4)
    assert LL[0, 0] == -np.inf


def test_reference_tolerance_mode(sol_gold):
    """BDF at the reference's reltol=abstol=1e-4: regression against the frozen values, and the
    size of its distance from the converged answer (~1e-3 relative)."""
    idx = np.arange(8)
    LL = co.llh_batch(sol_gold["X"][idx], [sol_gold["data"][:, :, 0]], [0.1], 0, 1e-4, 1e-4)
    assert relerr(LL[:, 0], sol_gold["LL_bdf1e4"][idx, 0]) < 1e-12
    rel = np.abs(LL[:, 0] - sol_gold["LL"][idx, 0]) / np.abs(sol_gold["LL"][idx, 0])
    assert 1e-6 < rel.max() < 5e-2


def test_redundant_per_patient_solves_give_same_numbers(sol_gold):
    datas = [sol_gold["data"][:, :, m] for m in range(4)]
    a = co.llh_batch(sol_gold["X"][:3], datas, [0.1] * 4, 0, 1e-4, 1e-4, redundant=False)
    b = co.llh_batch(sol_gold["X"][:3], datas, [0.1] * 4, 0, 1e-4, 1e-4, redundant=True)
    assert np.array_equal(a, b)


def test_python_llh_epilogue_matches_c(sol_gold):
    for i in (0, 7):
        for m in range(4):
            assert abs(o.llh_from_traj(sol_gold["traj"][i], sol_gold["data"][:, :, m], 0.1) - sol_gold["LL"][i, m]) \
                <= 1e-13 * abs(sol_gold["LL"][i, m])


def test_logprior_golden(sol_gold):
    lp = np.array([o.logprior(x, sol_gold["gmm_means"], sol_gold["gmm_stds"]) for x in sol_gold["X"]])
    assert np.array_equal(np.isfinite(lp), np.isfinite(sol_gold["logprior"]))
    f = np.isfinite(lp)
    assert relerr(lp[f], sol_gold["logprior"][f]) < 1e-13
    x = sol_gold["X"][0].copy()
    x[3] = 5.0001 * o.REFPARMS[3]       # outside the box (model.jl:51,70-71)
    assert o.logprior(x, sol_gold["gmm_means"], sol_gold["gmm_stds"]) == -np.inf


def test_em_golden_and_properties(sol_gold):
    L = sol_gold["em_L"]
    K, M = L.shape
    w = np.full(K, 1.0 / K)
    for it in range(1, 11):
        w = o.emiteration(w, L)
        if it == 1:
            assert relerr(w, sol_gold["em_w1"]) < 1e-14
    assert relerr(w, sol_gold["em_w10"]) < 1e-13
    assert relerr(co.emiteration(np.full(K, 1.0 / K), L), sol_gold["em_w1"]) < 1e-13
    assert relerr(co.emiteration(np.full(K, 1.0 / K), L, mt=True), sol_gold["em_w1"]) < 1e-13
    # (iii) sum(w) = 1 after one step for any positive w0
    rng = np.random.default_rng(0)
    assert abs(o.emiteration(rng.random(K) * 3, L).sum() - 1) < 1e-12
    # (iv) equal rows: w -> w / sum(w)
    Le = np.tile(rng.random((1, M)), (50, 1))
    w0 = rng.random(50)
    assert np.allclose(o.emiteration(w0, Le), w0 / w0.sum(), rtol=1e-13)
    # emiterations returns w0 first and niter entries (em.jl:5)
    ws = o.emiterations(np.full(K, 1.0 / K), L, 3)
    assert len(ws) == 3 and np.array_equal(ws[0], np.full(K, 1.0 / K)) and relerr(ws[1], sol_gold["em_w1"]) < 1e-14


def test_weightedchain_pieces():
    rng = np.random.default_rng(1)
    S = rng.random((6, 116))
    S[2] = S[1]
    S[3] = S[1]
    rows, counts = o.collapse_duplicates([S[:4], S[4:]])
    assert counts.tolist() == [1, 3, 1, 1] and rows.shape == (4, 116)
    LL = rng.standard_normal((4, 3)) * 5
    LL[2, 1] = -np.inf
    lhs, w, d = o.weightedchain_normalize(rng.standard_normal(4), LL, counts)
    assert np.allclose(lhs.sum(0), 1) and lhs[2, 1] == 0 and abs(w.sum() - 1) < 1e-15 and abs(d.sum() - 1) < 1e-15


def test_amm_oracle_step_properties():
    """Adaptive Metropolis restatement: the running moments equal the sample moments of the
    visited states, the adapted factor reproduces scale^2/d times their covariance, and before m > 2d only the
    initial proposal is used."""
    rng = np.random.default_rng(21)
    d = 6
    SL = np.eye(d) * 0.3
    x = rng.standard_normal(d)
    logf = lambda v: -0.5 * float(v @ v)
    tune = o.amm_new_tune(x, SL)
    cur = logf(x)
    visited = [x.copy()]
    for it in range(60):
        z, u, um = rng.standard_normal(d), rng.random(), rng.random()
        before = x.copy()
        x, cur, acc = o.amm_step(x, cur, tune, z, u, um, logf)
        if tune["m"] <= 2 * d:
            assert (not acc) or np.allclose(x, before + SL @ z)
        visited.append(x.copy())
    V = np.array(visited)
    assert tune["m"] == 60
    assert np.allclose(tune["Mv"], V.mean(0), rtol=1e-12)
    assert np.allclose(tune["Mvv"], (V[:, :, None] * V[:, None, :]).mean(0), rtol=1e-12)
    p = 60 / 61.0
    Sigma = (2.38 ** 2 / d / p) * np.cov(V.T, bias=True)
    Lm = tune["SigmaLm"]
    assert np.allclose(Lm @ Lm.T, Sigma, rtol=1e-9, atol=1e-12) and np.allclose(Lm, np.tril(Lm))
    assert o.chol_fullrank(np.ones((3, 3))) is None                      # rank 1
    A = rng.standard_normal((5, 5))
    S = A @ A.T + np.eye(5)
    assert np.allclose(o.chol_fullrank(S), np.linalg.cholesky(S), rtol=1e-12)


def test_every_golden_is_cross_checked_by_an_independent_integrator(sol_gold):
    """tests/golden/radau_crosscheck.json (written by tests/golden/crosscheck_radau.py): scipy's Radau IIA at 1e-12 on
    the restated RHS against the C oracle's extrapolation integrator, for EVERY finite sample of the fixtures."""
    import json
    r = json.load(open(os.path.join(G, "radau_crosscheck.json")))
    fin = np.flatnonzero(np.isfinite(sol_gold["LL"][:, 0]))
    got = {e["index"]: e for e in r["samples"] if e["llh_rel"] is not None}
    assert sorted(got) == fin.tolist() and r["n_checked"] == fin.size
    assert all(np.isfinite(e["llh_rel"]) and e["llh_rel"] < 1e-9 and e["traj_rel"] < 1e-8 for e in got.values())
    assert r["max_llh_rel"] < 1e-9 and r["max_traj_rel"] < 1e-8

