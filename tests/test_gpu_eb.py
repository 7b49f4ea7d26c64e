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
    # more entries per sample than the kernel's 1/sigma table holds (D = 600 > 512): the dividing path
    sig = rng.random((40, 15)) + 0.5
    xs = [rng.standard_normal((40, 15)) * 0.1 for _ in range(9)]
    ys = [rng.standard_normal((40, 15)) * 0.1 for _ in range(70)]
    xs[3][7, 2] = np.nan
    # (without renormalisation: the reference's sqrt(prod(sigma) (2 pi)^600) overflows to Inf in double precision; the
    # kernel works with its logarithm and does not)
    L = eb.likelihoodmat_nanfast(xs, ys, eb.MatrixNormalCentered(sig), renormalize=False)
    assert relerr(L, eo.likelihoodmat_nanfast(xs, ys, sig, renormalize=False)) < 1e-10
    assert np.all(np.isfinite(eb.likelihoodmat_nanfast(xs, ys, eb.MatrixNormalCentered(sig))))


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


def test_zsharded_model_two_shards_one_gpu(gpu):
    """dist.ShardedEB on one GPU: the z-matrix split into two row shards (the collectives emulated by combining the two
    shards' buffers) reproduces the unsharded model and the oracle: global renormalisation, hz, dhz, mple iterates."""
    import torch
    import eb_oracle as eo
    from gync_b200 import eb
    rng = np.random.default_rng(17)
    K, M, zm = 150, 9, 2
    base = np.array([40.0, 5.0, 150.0, 6.0])
    ys = [base * (1 + 0.03 * rng.standard_normal((31, 4))) for _ in range(K)]
    datas = patient_like(rng, M, 0.4)
    datas = [0.05 * d + 0.95 * base for d in datas]
    err = eb.MatrixNormalCentered(SIG * 0.3)
    zs = [y + err.rand(rng) for _ in range(zm) for y in ys]
    full = eb.LikelihoodModel(ys, ys, zs, datas, err)
    Ld, Lz = full.matrices()
    assert relerr(Lz, eo.likelihoodmat_nanfast(zs, ys, err.sigmas)) < 1e-9
    Y, Dm, Zs = eb._cols(ys), eb._cols(datas), eb._cols(zs)
    sig = np.ascontiguousarray(err.sigmas.ravel(order="F"))
    cut = 130
    # two "ranks" on one GPU through the C ABI; the collectives are emulated by combining the shards' buffers
    lib, L_ = gpu.lib.load(), gpu.lib
    import ctypes as C
    hs, scals = [], []
    for lo, hi in ((0, cut), (cut, K * zm)):
        h = C.c_void_p()
        sc = np.empty(2)
        L_.check(lib.gync_ebmodel_create_zshard(L_.dptr(Y), K, L_.dptr(Dm), M, L_.dptr(np.asfortranarray(Zs[:, lo:hi])), hi - lo, lo, K * zm,
                                                124, L_.dptr(sig), None, L_.dptr(sc), C.byref(h)))
        hs.append(h)
        scals.append(sc)
    g = np.array([min(scals[0][0], scals[1][0]), max(scals[0][1], scals[1][1])])
    for h in hs:
        L_.check(lib.gync_ebmodel_zshard_finish(h, L_.dptr(g)))
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    w = rng.random(K)
    w /= w.sum()
    dw = torch.from_numpy(w).to(dev)
    parts = [torch.empty(K, dtype=torch.float64, device=dev) for _ in hs]
    hzp = [torch.empty(1, dtype=torch.float64, device=dev) for _ in hs]
    for h, pt, hp in zip(hs, parts, hzp):
        L_.check(lib.gync_ebmodel_dhz_dev(h, dw.data_ptr(), 0, pt.data_ptr(), st))
        L_.check(lib.gync_ebmodel_hz_dev(h, dw.data_ptr(), hp.data_ptr(), st))
    torch.cuda.synchronize()
    d = (parts[0] + parts[1]).cpu().numpy()
    assert relerr(d, eb.dhz(full, w)) < 1e-12 and relerr(d, eo.dhzdiscr(w, Lz)) < 1e-11
    hzs = float((hzp[0] + hzp[1]).item())
    assert abs(hzs - eb.hz(full, w)) < 1e-12 * abs(hzs)
    # cont form: the constant -1 is added by one shard only
    for h, pt in zip(hs, parts):
        L_.check(lib.gync_ebmodel_dhz_dev(h, dw.data_ptr(), 1, pt.data_ptr(), st))
    torch.cuda.synchronize()
    assert relerr((parts[0] + parts[1]).cpu().numpy(), eb.dhz(full, w, cont=True)) < 1e-12
    # device-resident ascent loop with the partial gradients summed each iteration == unsharded mple
    reg, n = 0.5, 6
    hstep = eo.mple_stepsize(1.0, reg, M)
    ws = [torch.full((K,), 1.0 / K, dtype=torch.float64, device=dev) for _ in hs]
    gb = torch.empty(K, dtype=torch.float64, device=dev)
    for _ in range(n - 1):
        for h, wv, pt in zip(hs, ws, parts):
            L_.check(lib.gync_ebmodel_dhz_dev(h, wv.data_ptr(), 0, pt.data_ptr(), st))
        tot = parts[0] + parts[1]
        for h, wv in zip(hs, ws):
            L_.check(lib.gync_ebmodel_dlogl_dev(h, wv.data_ptr(), gb.data_ptr(), st))
            L_.check(lib.gync_ebmodel_ascent_dev(h, wv.data_ptr(), tot.data_ptr(), reg, gb.data_ptr(), 1 - reg, hstep, st))
    torch.cuda.synchronize()
    ref = eb.mple(full, reg, 1.0, n)[-1]
    assert torch.equal(ws[0], ws[1])                                   # replicated step: identical on every rank
    assert np.max(np.abs(ws[0].cpu().numpy() - ref)) < 1e-14
    assert lib.gync_ebmodel_nan_flag(hs[0], st) == 0
    for h in hs:
        lib.gync_ebmodel_destroy(h)
    full.close()


def test_eb_full_size_properties(gpu):
    """BASELINE-size runs checked through size-independent properties (no oracle can run at these sizes in seconds):
    likelihood matrix 1e5 x 40 — permutation equivariance (bit-exact), argument symmetry, range; regulariser on a
    1e4 x 1e4 z-matrix — the checksums  sum_k w_k dlogl_k = M  and  sum_k w_k dhz_k = hz(w) - 1  (from
    regularizers.jl:7-10, 30-45, 59-79), one-pass == two-pass dhz; projection at K = 1e6 — idempotent, on the simplex."""
    import os
    import torch
    from gync_b200 import eb, lib as L_
    lib = L_.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    I, J, D = 100_000, 40, 124
    sig = torch.tensor(SIG.ravel(order="F"), device=dev)
    A = torch.randn(D, I, dtype=torch.float64, device=dev, generator=gen) * 0.3          # sample-major (stride_k = I)
    B = torch.randn(J, D, dtype=torch.float64, device=dev, generator=gen) * 0.3          # column-major D x J
    B[torch.rand(J, D, device=dev, generator=gen) < 0.4] = float("nan")
    Lm = torch.empty(J, I, dtype=torch.float64, device=dev)
    L_.check(lib.gync_likelihoodmat_dev(A.data_ptr(), I, I, 1, B.data_ptr(), J, 1, D, D, sig.data_ptr(), 1, Lm.data_ptr(), st))
    assert torch.isfinite(Lm).all() and Lm.min() >= 0
    perm = torch.randperm(I, device=dev, generator=gen)
    Ap = A[:, perm].contiguous()
    Lp = torch.empty_like(Lm)
    L_.check(lib.gync_likelihoodmat_dev(Ap.data_ptr(), I, I, 1, B.data_ptr(), J, 1, D, D, sig.data_ptr(), 1, Lp.data_ptr(), st))
    assert torch.equal(Lp, Lm[:, perm])                                   # rows follow their samples bit-exactly
    Lt = torch.empty(I, J, dtype=torch.float64, device=dev)               # J x I column-major: arguments swapped
    L_.check(lib.gync_likelihoodmat_dev(B.data_ptr(), J, 1, D, A.data_ptr(), I, I, 1, D, sig.data_ptr(), 1, Lt.data_ptr(), st))
    torch.cuda.synchronize()
    assert float(((Lt.T - Lm).abs() / Lm.clamp_min(1e-300)).max()) < 1e-12
    del A, Ap, Lm, Lp, Lt
    # regulariser at K = Z = 1e4, M = 40 (both matrices built on the device)
    K, M = 10_000, 40
    rng = np.random.default_rng(23)
    Y = np.asfortranarray(rng.standard_normal((D, K)) * 0.3)
    Z = np.asfortranarray(Y + rng.standard_normal((D, K)) * 0.1)
    Dm = np.asfortranarray(rng.standard_normal((D, M)) * 0.3)
    Dm[rng.random((D, M)) < 0.4] = np.nan
    sg = np.ascontiguousarray(SIG.ravel(order="F"))
    import ctypes as C
    h = C.c_void_p()
    L_.check(lib.gync_ebmodel_create(L_.dptr(Y), K, L_.dptr(Dm), M, L_.dptr(Z), K, D, L_.dptr(sg), None, C.byref(h)))
    w = rng.random(K)
    w /= w.sum()
    d = np.empty(K)
    L_.check(lib.gync_ebmodel_dlogl(h, L_.dptr(w), L_.dptr(d)))
    assert abs(np.dot(w, d) - M) < 1e-10 * M
    hz = C.c_double()
    L_.check(lib.gync_ebmodel_hz(h, L_.dptr(w), None, C.byref(hz)))
    g1 = np.empty(K)
    L_.check(lib.gync_ebmodel_dhz(h, L_.dptr(w), 0, L_.dptr(g1)))
    assert abs(np.dot(w, g1) - (hz.value - 1.0)) < 1e-10 * max(1.0, abs(hz.value))
    os.environ["GYNC_DHZ_TWO_PASS"] = "1"
    try:
        g2 = np.empty(K)
        L_.check(lib.gync_ebmodel_dhz(h, L_.dptr(w), 0, L_.dptr(g2)))
    finally:
        del os.environ["GYNC_DHZ_TWO_PASS"]
    assert relerr(g1, g2) < 1e-12                                          # one HBM pass == the reference's two matvecs
    ws = w.copy()
    L_.check(lib.gync_ebmodel_mple(h, L_.dptr(ws), 0.5, 1e-7, 5, 0, None))
    assert abs(ws.sum() - 1) < 1e-12 and ws.min() >= 0
    lib.gync_ebmodel_destroy(h)
    # projection at K = 1e6
    y = rng.standard_normal(1_000_000) * 1e-5 + 1e-6
    p = eb.projectsimplex(y)
    assert abs(p.sum() - 1) < 1e-11 and p.min() >= 0
    assert np.max(np.abs(eb.projectsimplex(p) - p)) < 1e-16
    t = (y - p)[p > 0]
    assert np.ptp(t) < 1e-15                                               # p = max(y - t, 0) with ONE threshold
