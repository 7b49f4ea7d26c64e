"""GPU parity: the empirical-Bayes kernels (csrc/gync_eb.cuh) through the C ABI against oracle/eb_oracle.py
(SURVEY.md §8f rows 2-3).  Tolerances: the oracle follows the reference's binomial-formula sqeucdist_nan, whose own
cancellation error (~1e-12 absolute in the squared distance) bounds the agreement of the likelihood matrix; everything
downstream of a given matrix is compared at 1e-12 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def patient_like(rng, n, miss):
    out = []
    for _ in range(n):
        x = rng.standard_normal((31, 4)) * np.array([30.0, 3.0, 100.0, 4.0]) + np.array([40.0, 5.0, 150.0, 6.0])
        x[rng.random((31, 4)) < miss] = np.nan
        out.append(x)
    return out


SIG = np.tile(np.array([12.0, 1.0, 40.0, 1.5]), (31, 1))


def test_likelihoodmat_vs_oracle(gpu):
    import eb_oracle as eo
    from gync_b200 import eb
    rng = np.random.default_rng(11)
    d = eb.MatrixNormalCentered(SIG)
    for (I, J, mi, mj) in [(70, 40, 0.0, 0.5), (130, 65, 0.3, 0.3), (5, 3, 0.9, 0.0), (64, 64, 0.0, 0.0), (1, 1, 0.0, 0.2)]:
        xs = [x * 0.05 + np.array([40.0, 5.0, 150.0, 6.0]) * 0.95 for x in patient_like(rng, I, mi)]   # keep exp() in range
        ys = [x * 0.05 + np.array([40.0, 5.0, 150.0, 6.0]) * 0.95 for x in patient_like(rng, J, mj)]
        for renorm in (True, False):
            L = eb.likelihoodmat_nanfast(xs, ys, d, renormalize=renorm)
            ref = eo.likelihoodmat_nanfast(xs, ys, SIG, renormalize=renorm)
            assert L.shape == (I, J)
            assert relerr(L, ref) < 1e-10, (I, J, renorm, relerr(L, ref))
    # a fully missing sample: every joint mask is empty -> sq = 0, norm = 1 before renormalisation
    xs = [np.full((31, 4), np.nan)] + patient_like(rng, 3, 0.2)
    ys = patient_like(rng, 4, 0.2)
    L = eb.likelihoodmat_nanfast(xs, ys, d, renormalize=False)
    assert np.all(L[0] == 1.0)
    # non-uniform sigmas and a non-31x4 shape (generic D)
    sig = rng.random((7, 3)) + 0.5
    xs = [rng.standard_normal((7, 3)) for _ in range(33)]
    ys = [rng.standard_normal((7, 3)) for _ in range(21)]
    ys[2][1, 1] = np.nan
    L = eb.likelihoodmat_nanfast(xs, ys, eb.MatrixNormalCentered(sig))
    assert relerr(L, eo.likelihoodmat_nanfast(xs, ys, sig)) < 1e-11


def test_likelihoodmat_device_strides_and_symmetry(gpu):
    """The _dev entry with sample-major operands (element strides) equals the column-major host entry, and
    likelihoodmat(xs, ys) = likelihoodmat(ys, xs)' (what lets one K x M matrix serve em and dlogl)."""
    import torch
    from gync_b200 import eb, lib as L_
    lib = L_.load()
    rng = np.random.default_rng(12)
    I, J, D = 300, 77, 124
    A = rng.standard_normal((D, I)) * 0.3
    B = rng.standard_normal((D, J)) * 0.3
    A[rng.random(A.shape) < 0.2] = np.nan
    sig = np.ascontiguousarray(SIG.ravel(order="F"))
    ref = np.empty((I, J), order="F")
    L_.check(lib.gync_likelihoodmat(L_.dptr(np.asfortranarray(A)), I, L_.dptr(np.asfortranarray(B)), J, D, L_.dptr(sig), 1, L_.dptr(ref)))
    dev = torch.device("cuda", 0)
    dA = torch.from_numpy(np.ascontiguousarray(A)).to(dev)          # row-major D x I == sample-major (stride_k = I, stride_i = 1)
    dB = torch.from_numpy(np.ascontiguousarray(B.T)).to(dev)        # row-major J x D == column-major D x J (stride_k = 1, stride_j = D)
    dS = torch.from_numpy(sig).to(dev)
    dL = torch.empty(J, I, dtype=torch.float64, device=dev)         # == I x J column-major
    st = torch.cuda.current_stream().cuda_stream
    L_.check(lib.gync_likelihoodmat_dev(dA.data_ptr(), I, I, 1, dB.data_ptr(), J, 1, D, D, dS.data_ptr(), 1, dL.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(dL.cpu().numpy().T, ref)
    dLt = torch.empty(I, J, dtype=torch.float64, device=dev)        # == J x I column-major
    L_.check(lib.gync_likelihoodmat_dev(dB.data_ptr(), J, 1, D, dA.data_ptr(), I, I, 1, D, dS.data_ptr(), 1, dLt.data_ptr(), st))
    torch.cuda.synchronize()
    assert relerr(dLt.cpu().numpy(), ref) < 1e-13


def test_regularisers_vs_oracle(gpu):
    import eb_oracle as eo
    import gync_oracle as o
    from gync_b200 import eb
    rng = np.random.default_rng(13)
    for (K, M, zm) in [(50, 7, 1), (300, 40, 2), (1000, 130, 1), (257, 3, 3)]:
        Ld = rng.random((K, M)) + 1e-3
        Lz = rng.random((K * zm, K)) + 1e-3
        w = rng.random(K)
        w /= w.sum()
        m = eb.LikelihoodModel.from_matrices(Ld, Lz)
        assert abs(eb.logl(m, w) - eo.logl(w, Ld.T)) < 1e-12 * abs(eo.logl(w, Ld.T))
        assert relerr(eb.dlogl(m, w), eo.dlogl(w, Ld.T)) < 1e-12
        wz = eo.repeatweights(w, K * zm)
        assert abs(eb.hz(m, w) - eo.hz(w, Lz, wz)) < 1e-12 * abs(eo.hz(w, Lz, wz))
        wz2 = rng.random(K * zm)
        assert abs(eb.hz(m, w, wz2) - eo.hz(w, Lz, wz2)) < 1e-12 * abs(eo.hz(w, Lz, wz2))
        assert relerr(eb.dhz(m, w), eo.dhzdiscr(w, Lz)) < 1e-12
        assert relerr(eb.dhz(m, w, cont=True), eo.dhzloop2(Lz, w, wz)) < 1e-12
        ws = eb.em(m, w, 6)
        wo = o.emiterations(w, Ld, 6)
        assert len(ws) == 6 and np.array_equal(ws[0], w)
        assert max(relerr(a, b) for a, b in zip(ws, wo)) < 1e-12
        m.close()
    # hz skips rows with rhoz == 0 (regularizers.jl:41); dhz raises on NaN (:77)
    Lz = rng.random((20, 20))
    Lz[3, :] = 0.0
    m = eb.LikelihoodModel.from_matrices(rng.random((20, 4)), Lz)
    w = np.full(20, 0.05)
    assert abs(eb.hz(m, w) - eo.hz(w, Lz, w)) < 1e-13
    with pytest.raises(Exception, match="NaN in dhz"):
        eb.dhz(m, w)
    m.close()


def test_projectsimplex_vs_oracle(gpu):
    import eb_oracle as eo
    from gync_b200 import eb
    rng = np.random.default_rng(14)
    for n in (1, 2, 33, 1000, 1025, 200_000):
        for scale, shift in ((1.0, 0.0), (1e-3, 1.0 / n), (10.0, -3.0)):
            y = rng.standard_normal(n) * scale + shift
            p = eb.projectsimplex(y)
            ref = eo.projectsimplex_sort(y)
            assert np.max(np.abs(p - ref)) < 1e-14, (n, scale)
            assert abs(p.sum() - 1) < 1e-12 and np.all(p >= 0)
            assert np.array_equal(p == 0, ref == 0)
    y = np.array([0.2, 0.2, 0.2, 0.2, 0.2])                  # ties, already on the simplex
    assert np.allclose(eb.projectsimplex(y), y, atol=1e-16)
    y = np.array([5.0, 5.0, -1.0])                           # exact ties at the top
    assert np.allclose(eb.projectsimplex(y), [0.5, 0.5, 0.0])
    with pytest.raises(TypeError):
        eb.projectsimplex_bang([0.5, 0.5])


def test_mple_gradient_ascent_vs_oracle(gpu):
    import eb_oracle as eo
    from gync_b200 import eb
    rng = np.random.default_rng(15)
    K, M, zm = 400, 12, 2
    Ld = rng.random((K, M)) + 0.05
    Lz = rng.random((K * zm, K)) + 0.05
    m = eb.LikelihoodModel.from_matrices(Ld, Lz)
    w0 = eb.uniformweights(K)
    for reg in (0.0, 0.5, 1.0):
        n = 12
        h = eo.mple_stepsize(1.0, reg, M)
        ref = eo.gradientascent(eo.dmple_obj(Ld.T, Lz, reg), w0, n, h)
        got = eb.mple(m, reg, 1.0, n)
        assert len(got) == n and np.array_equal(got[0], w0)
        assert max(np.max(np.abs(a - b)) for a, b in zip(got, ref)) < 1e-13, reg
    # a step large enough to hit the boundary exercises movefromboundary (gradientascent.jl:19-25)
    ref = eo.gradientascent(eo.dmple_obj(Ld.T, Lz, 0.0), w0, 6, 5e-3)
    got = eb.mple(m, 0.0, 5e-3, 6, autoh=False)
    assert any(np.any(eo.projectsimplex(w + 5e-3 * eo.dlogl(w, Ld.T)) == 0) for w in ref[:-1])
    assert max(np.max(np.abs(a - b)) for a, b in zip(got, ref)) < 1e-13
    # the generic host-loop form with a device projection gives the same iterates
    got2 = eb.gradientascent(eb.dmple_obj(m, 0.0), w0, 6, 5e-3)
    assert max(np.max(np.abs(a - b)) for a, b in zip(got2, ref)) < 1e-13
    m.close()


def test_gyncmodel_pipeline(gpu):
    """gyncmodel (eb/gync.jl:10-25): samples -> phi on the device -> both likelihood matrices -> em / mple; checked
    against the oracle fed with the device trajectories, and phi against the golden trajectories."""
    import os
    import eb_oracle as eo
    import gync_oracle as o
    from gync_b200 import api, eb
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "solution_golden.npz"))
    rng = np.random.default_rng(16)
    x0 = o.init_x()
    X = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((48, 116)))
    X[0] = x0
    datas = api.alldatas()[:9]
    m = eb.gyncmodel(list(X), datas, zmult=2, sigma=0.1, seed=3)
    K = len(m.xs)
    assert K >= 40 and len(m.zs) == 2 * K and m.ys[0].shape == (31, 4)
    y_ref = api.forwardsol(x0)[:, api.measuredinds - 1]
    assert np.array_equal(m.ys[0], y_ref)                        # phi == forwardsol(x)[:, measuredinds]
    assert np.allclose(m.measerr.sigmas[0], [12.0, 1.0, 40.0, 1.5])
    Ld, Lz = m.matrices()
    assert relerr(Ld, eo.likelihoodmat_nanfast(m.ys, m.datas, m.measerr.sigmas)) < 1e-9
    assert relerr(Lz, eo.likelihoodmat_nanfast(m.zs, m.ys, m.measerr.sigmas)) < 1e-9
    w = eb.uniformweights(m)
    assert relerr(eb.dlogl(m, w), eo.dlogl(w, Ld.T)) < 1e-12
    assert relerr(eb.dhz(m, w), eo.dhzdiscr(w, Lz)) < 1e-12
    ws = eb.mple(m, 0.5, 1.0, 5)
    ref = eo.gradientascent(eo.dmple_obj(Ld.T, Lz, 0.5), w, 5, eo.mple_stepsize(1.0, 0.5, len(datas)))
    assert max(np.max(np.abs(a - b)) for a, b in zip(ws, ref)) < 1e-13
    m.close()
