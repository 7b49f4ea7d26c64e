"""CPU: the empirical-Bayes oracle (oracle/eb_oracle.py) against the reference's own consistency checks
(the reference holds no numeric vectors for these functions, test/priorestimation.jl:17-29)."""
import numpy as np
import pytest

import eb_oracle as eo
import gync_oracle as o


def _patient_like(rng, n, D=(31, 4), miss=0.4):
    out = []
    for _ in range(n):
        x = rng.standard_normal(D) * np.array([30.0, 3.0, 100.0, 4.0]) + np.array([40.0, 5.0, 150.0, 6.0])
        x[rng.random(D) < miss] = np.nan
        out.append(x)
    return out


def test_sqeucdist_nan_equals_its_definition():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((124, 17))
    B = rng.standard_normal((124, 9))
    A[rng.random(A.shape) < 0.3] = np.nan
    B[rng.random(B.shape) < 0.5] = np.nan
    a, b = eo.sqeucdist_nan(A, B), eo.sqeucdist_nan_direct(A, B)
    assert np.max(np.abs(a - b)) < 1e-11
    # no NaNs: plain squared distances
    A, B = rng.standard_normal((5, 4)), rng.standard_normal((5, 3))
    ref = ((A[:, :, None] - B[:, None, :]) ** 2).sum(0)
    assert np.allclose(eo.sqeucdist_nan(A, B), ref, rtol=0, atol=1e-13)


def test_likelihoodmat_nanfast_vs_elementwise_pdf():
    """likelihoodmat.jl:9 (the commented element-wise form pdf(d, z-y)) vs the fast form without renormalisation."""
    rng = np.random.default_rng(2)
    sig = np.tile(np.array([12.0, 1.0, 40.0, 1.5]), (31, 1))
    xs, ys = _patient_like(rng, 6), _patient_like(rng, 5, miss=0.0)
    L = eo.likelihoodmat_nanfast(xs, ys, sig, renormalize=False)
    for i, x in enumerate(xs):
        for j, y in enumerate(ys):
            d = x - y
            ok = ~np.isnan(d)
            # pdf = exp(-sq/2) / prod(sigma sqrt(2 pi)) over the observed entries (utils.jl:93-101)
            lognorm = np.sum(np.log(sig[ok])) + 0.5 * ok.sum() * np.log(2 * np.pi)
            assert np.isclose(np.log(L[i, j]), eo.matrixnormal_logpdf(sig, d) + lognorm, rtol=0, atol=1e-9)
    # renormalised: the largest possible entry is 1 only if the closest pair also has the largest normalisation;
    # what always holds: entries are positive, finite, and the matrix is symmetric in its arguments
    Lr = eo.likelihoodmat_nanfast(xs, ys, sig)
    assert np.all(np.isfinite(Lr)) and np.all(Lr >= 0)
    assert np.allclose(Lr, eo.likelihoodmat_nanfast(ys, xs, sig).T, rtol=1e-12, atol=0)


def test_dhz_forms_agree_and_match_gradient_of_hz():
    """dhztest (regularizers.jl:157-170): dhzmatrix == dhzloop == dhzloop2 == d hz / d w."""
    rng = np.random.default_rng(3)
    n, m = 40, 30
    L, wz, w = rng.random((n, m)), rng.random(n), rng.random(m)
    a, b, c = eo.dhzmatrix(L, w, wz), eo.dhzloop(L, w, wz), eo.dhzloop2(L, w, wz)
    assert np.allclose(a, b, rtol=1e-12) and np.allclose(a, c, rtol=1e-12)
    # the "-1" in dhz* stands for the dependence of wz on w; with wz held fixed the gradient of hz is
    # -L' (wz ./ rhoz), the first term of dhzdiscr (regularizers.jl:70-72)
    g = np.empty(m)
    for k in range(m):
        e = np.zeros(m)
        e[k] = 1e-6
        g[k] = (eo.hz(w + e, L, wz) - eo.hz(w - e, L, wz)) / 2e-6
    assert np.allclose(g, -(L.T @ (wz / (L @ w))), rtol=1e-7)
    # dhzdiscr on a zmult = 2 model: total derivative of hz(w, L, repeatweights(w))
    K, zm = 12, 2
    L2 = rng.random((K * zm, K)) + 0.1
    w2 = rng.random(K)
    w2 /= w2.sum()
    d = eo.dhzdiscr(w2, L2)
    g2 = np.empty(K)
    for k in range(K):
        e = np.zeros(K)
        e[k] = 1e-7
        g2[k] = (eo.hz(w2 + e, L2, eo.repeatweights(w2 + e, K * zm)) - eo.hz(w2 - e, L2, eo.repeatweights(w2 - e, K * zm))) / 2e-7
    assert np.allclose(d, g2, rtol=1e-6)


def test_projectsimplex_heap_equals_sort_and_is_the_projection():
    rng = np.random.default_rng(4)
    for n in (1, 2, 7, 100, 1000):
        for scale in (0.01, 1.0, 10.0):
            y = rng.standard_normal(n) * scale
            a, b = eo.projectsimplex_heap(y), eo.projectsimplex_sort(y)
            assert np.array_equal(a, b)
            assert abs(a.sum() - 1) < 1e-12 and np.all(a >= 0)
            # KKT: y - a is constant on the support and >= that constant off it... i.e. a = max(y - t, 0)
            t = (y - a)[a > 0]
            assert np.ptp(t) < 1e-12
    w = np.full(10, 0.1)
    assert np.allclose(eo.projectsimplex(w), w)                     # already on the simplex


def test_dlogl_is_the_em_update_and_logl_increases():
    """em.jl:27-41 vs regularizers.jl:7-10: emiteration!(w, L) = w .* dlogl(w, L') / M; EM ascends logl."""
    rng = np.random.default_rng(5)
    K, M = 200, 7
    L = rng.random((K, M))
    w = rng.random(K)
    w /= w.sum()
    assert np.allclose(o.emiteration(w, L), w * eo.dlogl(w, L.T) / M, rtol=1e-13)
    w1 = o.emiteration(w, L)
    assert eo.logl(w1, L.T) > eo.logl(w, L.T)
    # gradient of logl
    g = np.empty(K)
    for k in range(K):
        e = np.zeros(K)
        e[k] = 1e-7
        g[k] = (eo.logl(w + e, L.T) - eo.logl(w - e, L.T)) / 2e-7
    assert np.allclose(g, eo.dlogl(w, L.T), rtol=1e-6)


def test_gradientascent_stays_on_simplex_and_ascends():
    rng = np.random.default_rng(6)
    K, M, zm = 30, 5, 2
    Ld = rng.random((K, M)) + 0.05
    Lz = rng.random((K * zm, K)) + 0.05
    w0 = np.full(K, 1 / K)
    h = eo.mple_stepsize(1.0, 0.5, M)
    ws = eo.gradientascent(eo.dmple_obj(Ld.T, Lz, 0.5), w0, 20, h)
    assert len(ws) == 20 and np.array_equal(ws[0], w0)
    obj = [0.5 * eo.hz(w, Lz, eo.repeatweights(w, K * zm)) + 0.5 * eo.logl(w, Ld.T) for w in ws]
    assert all(abs(w.sum() - 1) < 1e-12 and np.all(w >= 0) for w in ws)
    assert obj[-1] > obj[0]
    assert np.array_equal(eo.movefromboundary(np.array([0.5, 0.5]), np.array([1.0, 0.0])), np.array([0.75, 0.25]))


def test_psi_S_preserves_the_simplex():
    rng = np.random.default_rng(7)
    w = rng.random(50)
    w /= w.sum()
    assert abs(eo.psi_S(w, rng.standard_normal(50) * 0.1).sum() - 1) < 1e-13
