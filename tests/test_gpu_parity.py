"""GPU suite: the CUDA path (through the C ABI) against the oracle, the frozen golden vectors, and
size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): trajectories at the measurement times within rtol 1e-6,
log-likelihoods within 1e-8 relative, EM weights within 1e-10 after a fixed number of iterations —
against the CONVERGED solution of the reference's ODE (the reference's own CVODE run
at reltol=abstol=1e-4 is ~1e-3 away from it, and CVODE is not available to clone step for step).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
RTOL_TRAJ, RTOL_LLH, RTOL_EM = 1e-6, 1e-8, 1e-10


def relerr(a, b):
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "solution_golden.npz"))


def near_set(n, seed):
    import gync_oracle as o
    rng = np.random.Generator(np.random.PCG64(seed))
    X = np.exp(np.log(o.init_x())[None, :] + 0.0998 * rng.standard_normal((n, 116)))
    return X


# ---------------------------------------------------------------------------------------- K1
def test_loglik_and_trajectories_vs_golden(gpu, gold):
    datas = [gold["data"][:, :, m] for m in range(4)]
    LL, st, Ym = gpu.loglik_matrix(gold["X"], datas, [0.1] * 4, want_traj=True)
    ok = np.isfinite(gold["LL"][:, 0])
    assert np.array_equal(st == 0, ok)                         # failures: same samples as the oracle
    assert np.all(LL[~ok] == -np.inf) and np.isnan(Ym[~ok]).all()   # gyncycle.jl:18-19, model.jl:138-141
    assert relerr(Ym[ok], gold["traj"][ok]) < RTOL_TRAJ
    assert relerr(LL[ok], gold["LL"][ok]) < RTOL_LLH


def test_loglik_vs_c_oracle_fresh_samples(gpu):
    """Seeded samples that are not in the fixtures; the C oracle (converged mode) runs here in seconds."""
    import c_oracle as co
    import gync_oracle as o
    X = near_set(48, 777)
    X[5, 115] = 27.0            # integer period -> duplicated times
    X[6, 115] = 31.75
    data = o.lausanne(2)
    LLo, tro = co.llh_batch(X, [data], [0.15], 1, 1e-12, 1e-12, want_traj=True)
    LL, st, Ym = gpu.loglik_matrix(X, [data], [0.15], want_traj=True)
    assert np.all(st == 0) and np.all(np.isfinite(LLo))
    assert relerr(Ym, tro) < RTOL_TRAJ
    assert relerr(LL, LLo) < RTOL_LLH


def test_parity_tail_every_case_within_bound(gpu):
    """The north-star bound holds for EVERY case, not for a hand-picked hundred: 30 720 near-set (sample, patient) cases
    (three seeded populations x 3 Lausanne patients each; the third was drawn after every knob of the solver had been
    fixed on the first two) against the frozen converged log-likelihoods of the C oracle
    (tests/golden/nearset_llh_golden.npz, oracle at 1e-12), at the library default tolerance.  Round 1 (RODAS4 at
    1e-8/1e-10) had 0.08 % of these above 1e-8 (max 4.6e-8); see DESIGN.md section 3 for the mechanism that removed them."""
    import sys
    sys.path.insert(0, G)
    import gync_oracle as o
    import make_nearset_golden as mg
    g = np.load(os.path.join(G, "nearset_llh_golden.npz"))
    la = o.patients()[0]
    stats = {}
    for name in ("A", "B", "C"):
        seed, n, pats = mg.SETS[name]
        X = mg.near_set(seed, n)
        LL, st = gpu.loglik_matrix(X, [la[i] for i in pats], [0.1] * len(pats))
        assert np.all(st == 0)
        E = np.abs(LL - g["LL_" + name]) / np.abs(g["LL_" + name])
        stats[name] = (float(np.median(E)), float(np.quantile(E, 0.99)), float(E.max()), float((E > RTOL_LLH).mean()))
        assert E.max() <= RTOL_LLH, (name, stats[name], np.argwhere(E > RTOL_LLH)[:8])
    print("parity tail (median, p99, max, frac>1e-8):", stats)


def test_wide_set_failure_set_and_values(gpu):
    """The wide set of eb/gync.jl:59 (parameters uniform on (0, 5 ref)): 77 % of the draws do not solve.  The failure SET
    must equal the oracle's, failed samples give -Inf, the solvable ones obey the same 1e-8 bound."""
    import sys
    sys.path.insert(0, G)
    import gync_oracle as o
    import make_nearset_golden as mg
    g = np.load(os.path.join(G, "nearset_llh_golden.npz"))
    Xw = mg.wide_set(int(g["seed_W"]), 512)
    LL, st = gpu.loglik_matrix(Xw, [o.patients()[0][0]], [0.1])
    fin = np.isfinite(g["LL_W"])
    assert 0.1 < fin.mean() < 0.5
    assert np.array_equal(st == 0, fin), np.flatnonzero((st == 0) != fin)
    assert np.all(LL[~fin, 0] == -np.inf)
    # the oracle's own conditioning of each fixture value: its 1e-12 and 1e-13 runs (tests/golden/add_wide_conditioning.py)
    with np.errstate(invalid="ignore"):
        unc = np.abs(g["LL_W_1e13"] - g["LL_W"]) / np.abs(g["LL_W"])
        E = np.abs(LL[:, 0] - g["LL_W"]) / np.abs(g["LL_W"])
    well = fin & (unc <= 1e-10)
    assert well.sum() >= 100
    assert E[well].max() <= RTOL_LLH, E[well].max()
    # 12 of the 117 solvable draws amplify integration error so much that the ORACLE moves by 1e-10 .. 1.1e-8 between its
    # two tightest settings; there the fixture cannot carry a 1e-8 bound, the check is 3e-8
    assert E[fin & ~well].max() <= 3e-8, E[fin & ~well].max()


def test_loglik_epilogue_in_isolation(gpu, gold):
    """Oracle likelihood of the KERNEL's own trajectory: isolates the fused epilogue (<=1e-13)."""
    import gync_oracle as o
    datas = [gold["data"][:, :, m] for m in range(4)]
    LL, st, Ym = gpu.loglik_matrix(gold["X"][:8], datas, [0.1, 0.2, 0.05, 0.1], want_traj=True)
    for i in range(8):
        for m, s in enumerate([0.1, 0.2, 0.05, 0.1]):
            ref = o.llh_from_traj(Ym[i], datas[m], s)
            assert abs(LL[i, m] - ref) <= 1e-13 * abs(ref)


def test_loglik_many_patients_path(gpu, gold):
    """M > 4 switches to solve-once + separate (sample x patient) scoring kernel: same numbers as the fused
    single-patient path, same failure semantics."""
    import gync_oracle as o
    la = o.patients()[0]
    datas = [gold["data"][:, :, m] for m in range(4)] + [la[1], la[4], np.full((31, 4), np.nan)]
    sig = [0.1, 0.2, 0.05, 0.1, 0.1, 0.3, 0.1]
    X = gold["X"]
    LL, st, Ym = gpu.loglik_matrix(X, datas, sig, want_traj=True)
    ok = st == 0
    assert np.array_equal(ok, np.isfinite(gold["LL"][:, 0]))
    assert np.all(LL[~ok] == -np.inf) and np.all(LL[:, 6] == -np.inf)          # failed solves; all-missing table
    for m in range(6):
        one, _ = gpu.loglik_matrix(X, [datas[m]], [sig[m]])                    # fused path
        assert relerr(LL[ok, m], one[ok, 0]) < 1e-13
    for i in np.flatnonzero(ok)[:6]:
        for m in range(6):
            ref = o.llh_from_traj(Ym[i], datas[m], sig[m])
            assert abs(LL[i, m] - ref) <= 1e-13 * abs(ref)
    assert relerr(LL[ok][:, [0, 3]], gold["LL"][ok][:, [0, 3]]) < RTOL_LLH     # sigma = 0.1 as in the fixture for m = 0, 3


def test_loglik_edge_cases(gpu, gold):
    x = gold["X"][:3].copy()
    nanpat = np.full((31, 4), np.nan)
    LL, st = gpu.loglik_matrix(x, [nanpat, gold["data"][:, :, 0]], [0.1, 0.1])
    assert np.all(LL[:, 0] == -np.inf) and np.all(np.isfinite(LL[:, 1]))        # all-missing table (model.jl:161)
    x[1, 115] = np.nan
    x[2, 85] = -1.0      # LH_R(0) < 0: negative base of the real power
    LL, st = gpu.loglik_matrix(x, [gold["data"][:, :, 0]], [0.1])
    assert st[0] == 0 and st[1] == 4 and st[2] != 0 and LL[1, 0] == -np.inf and LL[2, 0] == -np.inf
    # empty batch
    LL, st = gpu.loglik_matrix(np.empty((0, 116)), [gold["data"][:, :, 0]], [0.1])
    assert LL.shape == (0, 1)
    # step budget exhaustion is a per-sample status, not an error
    lib = gpu.lib.load()
    X = gpu.lib.fcol(gold["X"][:2])
    D = np.asfortranarray(gold["data"][:, :, :1])
    out = np.empty((2, 1), order="F")
    stt = np.zeros(2, dtype=np.int32)
    o_ = gpu.lib.opts(1e-8, 1e-8, max_steps=50)
    gpu.lib.check(lib.gync_loglik_batch(gpu.lib.dptr(X), 2, gpu.lib.dptr(D), 1, gpu.lib.dptr(np.array([0.1])), C.byref(o_),
                                        gpu.lib.dptr(out), None, stt.ctypes.data_as(gpu.lib.c_int32_p), None))
    assert np.all(stt == 1) and np.all(out == -np.inf)
    # bad arguments -> error code + message, no exception across the ABI
    rc = lib.gync_loglik_batch(None, 5, None, 1, None, None, None, None, None, None)
    assert rc == -1 and b"bad arguments" in lib.gync_last_error()


def test_api_llh_config_and_sampling(gpu, gold):
    c = gpu.Config()
    x0 = gpu.api.init(c)
    l = gpu.llh(c, x0)
    assert abs(l - gold["LL"][0, 0]) <= RTOL_LLH * abs(l)
    s = gpu.Sampling(gold["X"][:5], c, None)
    assert relerr(gpu.llh(s), gold["LL"][:5, 0]) < RTOL_LLH                       # llh(s::Sampling), sampling.jl:32


def test_gync_and_forwardsol_generic_grid(gpu, gold):
    import c_oracle as co
    import gync_oracle as o
    t = np.arange(0.0, 31.0)
    Y = gpu.gync(o.REFY0, o.allparms(o.REFPARMS), t)
    ref = co.gync(o.REFY0, o.allparms(o.REFPARMS), t, mode=1, rtol=1e-12, atol=1e-12)
    big = np.abs(ref) > 1e-6 * np.abs(ref).max(0)
    assert np.array_equal(Y[0], o.REFY0)                       # t[0] is the initial time (cvode semantics)
    assert np.max(np.abs(Y - ref)[big] / np.abs(ref)[big]) < RTOL_TRAJ
    Yf = gpu.forwardsol(gold["X"][:4], np.array([1.0, 2.5, 2.5, 17.0, 40.0]))
    for i in range(4):
        parms, y0, _ = o.split(gold["X"][i])
        ref = co.gync(y0, o.allparms(parms), np.array([1.0, 2.5, 17.0, 40.0]), mode=1, rtol=1e-12, atol=1e-12)
        got = Yf[i][[0, 1, 3, 4]][:, o.MEASUREDINDS]
        assert relerr(got, ref[:, o.MEASUREDINDS]) < RTOL_TRAJ
        assert np.array_equal(Yf[i][1], Yf[i][2])
    assert relerr(gpu.referencesolution()[:, o.MEASUREDINDS], gold["refsol"][:, o.MEASUREDINDS]) < RTOL_TRAJ
    bad = np.flatnonzero(~np.isfinite(gold["LL"][:, 0]))[0]
    assert np.isnan(gpu.forwardsol(gold["X"][bad])).all()      # failure -> NaN matrix


def test_full_size_properties(gpu, gold):
    """cfg 2 size (1e5 samples): order-invariance (bit-exact: persistent lanes must not change any
    sample's arithmetic), M-consistency, finite fraction."""
    N = 100_000
    X = near_set(N, 20160913)
    data = gold["data"][:, :, 0]
    LL1, st1 = gpu.loglik_matrix(X, [data], [0.1])
    assert (st1 == 0).mean() > 0.999
    perm = np.random.default_rng(5).permutation(N)
    LL2, st2 = gpu.loglik_matrix(X[perm], [data], [0.1])
    assert np.array_equal(LL2[:, 0], LL1[perm, 0]) and np.array_equal(st2, st1[perm])
    sub = slice(0, 4096)
    datas = [gold["data"][:, :, m] for m in range(4)]
    LL4, _ = gpu.loglik_matrix(X[sub], datas, [0.1] * 4)
    assert np.array_equal(LL4[:, 0], LL1[sub, 0])              # column m is independent of the other patients
    import c_oracle as co
    pick = np.random.default_rng(6).choice(N, 16, replace=False)
    LLo = co.llh_batch(X[pick], [data], [0.1], 1, 1e-12, 1e-12)
    assert relerr(LL1[pick], LLo) < RTOL_LLH


# ---------------------------------------------------------------------------------------- K2
def test_logprior_vs_golden(gpu, gold):
    c = gpu.Config(priory0=gpu.api.GaussianMixture(gold["gmm_means"], gold["gmm_stds"]))
    lp = gpu.logprior(c, gold["X"])
    f = np.isfinite(gold["logprior"])
    assert np.array_equal(np.isfinite(lp), f)
    assert relerr(lp[f], gold["logprior"][f]) < 1e-12
    x = gold["X"][0].copy()
    x[3] = 5.0001 * gpu.api.refparms[3]
    assert gpu.logprior(c, x) == -np.inf
    # the default prior is built from the device's own converged reference solution
    c2 = gpu.Config()
    assert abs(gpu.logprior(c2, gold["X"][0]) - gold["logprior"][0]) < 1e-6 * abs(gold["logprior"][0])


# ---------------------------------------------------------------------------------------- K4
def test_em_vs_golden_and_oracle(gpu, gold):
    import gync_oracle as o
    L = gold["em_L"]
    K, M = L.shape
    w = np.full(K, 1.0 / K)
    gpu.emiteration_bang(w, L, niter=1)
    assert relerr(w, gold["em_w1"]) < RTOL_EM
    gpu.emiteration_bang(w, L, niter=9)
    assert relerr(w, gold["em_w10"]) < RTOL_EM
    gpu.emiteration_bang(w, L, niter=90)
    assert relerr(w, gold["em_w100"]) < RTOL_EM
    ws = gpu.emiterations(np.full(K, 1.0 / K), L, 11)                       # em.jl:5: w0 first
    assert len(ws) == 11 and relerr(ws[1], gold["em_w1"]) < RTOL_EM and relerr(ws[10], gold["em_w10"]) < RTOL_EM
    rng = np.random.default_rng(2)
    for K, M in ((1, 1), (7, 3), (257, 8), (1000, 17), (5000, 40), (3001, 64), (70001, 33)):
        L = rng.random((K, M)) ** 4
        w0 = rng.random(K)
        ref = w0.copy()
        for _ in range(3):
            ref = o.emiteration(ref, L)
        got = gpu.emiteration_bang(w0.copy(), L, niter=3)
        assert relerr(got, ref) < RTOL_EM, (K, M)
    wc = gpu.WeightedChain(rng.random((100, 3)), rng.random((100, 5)), rng.random(100))    # test/priorestimation.jl:6-7
    before = wc.weights.copy()
    gpu.emiteration_bang(wc)
    assert relerr(wc.weights, o.emiteration(before, wc.likelihoods)) < RTOL_EM


def test_em_edge_and_full_size(gpu):
    rng = np.random.default_rng(3)
    # zero rows (samples whose likelihood underflowed) stay exactly zero; Inf/NaN propagate like the reference
    L = rng.random((1000, 5))
    L[17] = 0.0
    w = gpu.emiteration_bang(np.full(1000, 1e-3), L, niter=4)
    assert w[17] == 0.0 and abs(w.sum() - 1) < 1e-12
    L[:, 2] = 0.0                                                # norms[m] == 0 -> 0/0 (em.jl has no guard)
    assert np.isnan(gpu.emiteration_bang(np.full(1000, 1e-3), L)).all()
    with pytest.raises(ValueError):
        gpu.emiteration_bang(np.ones(3), np.ones((4, 2)))
    # cfg 5 size on one GPU: K = 1e6 x 40; properties (iii) sum w = 1, (iv) equal rows -> w/sum(w)
    K, M = 1_000_000, 40
    L = rng.random((K, M))
    w = gpu.emiteration_bang(rng.random(K), L)
    assert abs(w.sum() - 1) < 1e-11
    Le = np.tile(rng.random((1, M)), (K, 1))
    w0 = rng.random(K)
    assert relerr(gpu.emiteration_bang(w0.copy(), Le), w0 / w0.sum()) < 1e-11


def test_em_partly_resident_shard(gpu):
    """Rows-per-block beyond what shared memory holds (cfg 5's 8-GPU shard: 125 000 x 40 per GPU): the kernel keeps the rows
    that fit and streams the rest.  Result and trajectory equal the oracle's; with GYNC_EM_NO_HYBRID the all-streaming kernel
    gives the same weights to rounding."""
    import ctypes, gync_oracle as o
    libc = ctypes.CDLL(None)
    rng = np.random.default_rng(12)
    for K, M in ((125_003, 40), (118_001, 37), (140_000, 12)):
        Lm = rng.random((K, M))
        Lm[rng.integers(0, K, 50)] = 0.0
        w0 = rng.random(K)
        w0 /= w0.sum()
        ref = [w0]
        for _ in range(4):
            ref.append(o.emiteration(ref[-1], Lm))
        traj = gpu.emiterations(w0, Lm, 5)                        # w0 first, like em.jl:5
        for a, b in zip(traj, ref):
            assert relerr(np.asarray(a), b) < RTOL_EM
        w = gpu.emiteration_bang(w0.copy(), Lm, niter=4)
        libc.setenv(b"GYNC_EM_NO_HYBRID", b"1", 1)
        try:
            ws = gpu.emiteration_bang(w0.copy(), Lm, niter=4)
        finally:
            libc.unsetenv(b"GYNC_EM_NO_HYBRID")
        assert relerr(w, ref[-1]) < RTOL_EM and relerr(ws, w) < 1e-12 and abs(w.sum() - 1) < 1e-12


def test_em_sharded_step_kernels(gpu):
    """The kernels behind the multi-GPU path (norms / step with external denominators), emulated
    on one device as two shards with the 'allreduce' done on the host."""
    import torch
    import gync_oracle as o
    from gync_b200.dist import DeviceEM, shard_bounds, em_iterate_sharded
    rng = np.random.default_rng(4)
    K, M, niter = 20011, 40, 6
    L = rng.random((K, M))
    L /= L.sum(0, keepdims=True)
    w0 = np.full(K, 1.0 / K)
    shards = []
    for r in range(2):
        lo, hi = shard_bounds(K, 2, r)
        shards.append(DeviceEM(torch.from_numpy(w0[lo:hi].copy()).cuda(),
                               torch.from_numpy(np.ascontiguousarray(L[lo:hi].T)).cuda()))
    norms = shards[0].norms() + shards[1].norms()
    for _ in range(niter):
        norms = shards[0].step(norms) + shards[1].step(norms)
    ref = w0.copy()
    for _ in range(niter):
        ref = o.emiteration(ref, L)
    got = np.concatenate([s.w.cpu().numpy() for s in shards])
    assert relerr(got, ref) < RTOL_EM


def test_em_fused_kernel_world1(gpu):
    """The fused compute+exchange kernel with a single rank (mailbox = local memory): same numbers
    as the oracle.  The multi-rank exchange is exercised by tests/manual/multigpu_check.py under torchrun."""
    import torch
    import gync_oracle as o
    from gync_b200.dist import FusedEM
    rng = np.random.default_rng(12)
    for K, M in ((20011, 40), (513, 7), (100_000, 33)):
        L = rng.random((K, M))
        L /= L.sum(0, keepdims=True)
        w0 = rng.random(K)
        w0 /= w0.sum()
        em = FusedEM(torch.from_numpy(w0.copy()).cuda(), torch.from_numpy(np.ascontiguousarray(L.T)).cuda())
        em.iterate(3)
        em.iterate(2)                      # sequence numbers carry over between launches
        torch.cuda.synchronize()
        assert em.error() == 0
        ref = w0.copy()
        for _ in range(5):
            ref = o.emiteration(ref, L)
        assert relerr(em.w.cpu().numpy(), ref) < RTOL_EM
        em.close()


# ---------------------------------------------------------------------------------------- K3 / K5
def test_mcmc_step_with_injected_randomness(gpu, gold):
    """Deterministic parity of the Metropolis step: same normals/uniforms into the
    oracle's step function (logf through the converged C oracle) and into the kernel."""
    import c_oracle as co
    import gync_oracle as o
    lib, L = gpu.lib.load(), gpu.lib
    c = gpu.Config(priory0=gpu.api.GaussianMixture(gold["gmm_means"], gold["gmm_stds"]))
    Cn, iters = 4, 5
    rng = np.random.default_rng(8)
    z = rng.standard_normal((iters, 116, Cn))
    u = rng.random((iters, Cn))
    u[1] = 1e-300                      # force an accept
    u[3] = 1.0                         # force a reject
    datas = [gold["data"][:, :, m] for m in range(4)]
    sd = np.sqrt(np.log(1.01))
    chol = np.linalg.cholesky(np.eye(116) * np.log(1.01) + 1e-4 * np.ones((116, 116)))     # dense proposal for chain test

    def logf(xl, m):
        x = np.exp(xl)
        lp = o.logprior(x, gold["gmm_means"], gold["gmm_stds"])
        if lp == -np.inf:
            return lp
        return lp + co.llh_batch(x[None, :], [datas[m]], [0.1], 1, 1e-12, 1e-12)[0, 0] + xl.sum()

    for use_chol in (False, True):
        state = L.fcol(np.tile(np.log(o.init_x()), (Cn, 1)))
        lf = np.full(Cn, np.nan)
        out = np.empty((iters, 116, Cn), order="F")
        nacc = np.zeros(Cn, dtype=np.int64)
        pr = c.prior()
        op = L.opts(1e-9, 1e-9)
        D = np.asfortranarray(gold["data"])
        sig = np.full(4, 0.1)
        pat = np.arange(Cn, dtype=np.int32)
        L.check(lib.gync_mcmc_run(L.dptr(state), L.dptr(lf), Cn, L.dptr(L.fcol(chol)) if use_chol else None,
                                  0.0 if use_chol else sd, L.dptr(D), 4, L.dptr(sig), pat.ctypes.data_as(L.c_int32_p),
                                  C.byref(pr), C.byref(op), iters, 1, 0, 0, L.dptr(np.asfortranarray(z)),
                                  L.dptr(np.asfortranarray(u)), L.dptr(out), nacc.ctypes.data_as(L.c_int64_p)))
        for cidx in range(Cn):
            xs = np.log(o.init_x())
            cur = logf(xs, cidx)
            acc = 0
            SL = chol if use_chol else np.eye(116) * sd
            for it in range(iters):
                xs, cur, a = o.metropolis_step(xs, cur, SL, z[it, :, cidx], u[it, cidx], lambda v: logf(v, cidx))
                acc += a
                assert relerr(out[it, :, cidx], np.exp(xs)) < 1e-12          # rows are exp(state) (sampling.jl:74)
            assert acc == nacc[cidx] and acc >= 1
            assert abs(lf[cidx] - cur) <= 1e-8 * abs(cur)
            assert np.allclose(state[cidx], xs, rtol=0, atol=1e-13)


def test_amm_adaptive_step_with_injected_randomness(gpu, gold):
    """Config.adapt = true (Mamba AMM with adaptation): the kernel pair against the oracle's
    amm_step with the same normals / uniforms / mixture uniforms, started from a tuning state past m > 2d so that
    adapted proposals, the moment updates and the re-factorisation are all exercised, and from a fresh variate."""
    import c_oracle as co
    import gync_oracle as o
    lib, L = gpu.lib.load(), gpu.lib
    c = gpu.Config(priory0=gpu.api.GaussianMixture(gold["gmm_means"], gold["gmm_stds"]))
    Cn, iters = 3, 6
    rng = np.random.default_rng(18)
    datas = [gold["data"][:, :, m] for m in range(4)]
    sd = np.sqrt(np.log(1.01))
    x0 = np.log(o.init_x())

    def logf(xl, m):
        x = np.exp(xl)
        lp = o.logprior(x, gold["gmm_means"], gold["gmm_stds"])
        if lp == -np.inf:
            return lp
        return lp + co.llh_batch(x[None, :], [datas[m]], [0.1], 1, 1e-12, 1e-12)[0, 0] + xl.sum()

    def history_tune(seed, m):
        """A tuning state as m iterations of a slowly moving chain would leave it."""
        r = np.random.default_rng(seed)
        V = x0[None, :] + np.cumsum(r.standard_normal((m + 1, 116)) * 0.004, axis=0)
        t = o.amm_new_tune(V[0], np.eye(116) * sd)
        for k in range(1, m + 1):
            p = k / (k + 1.0)
            t["Mv"] = p * t["Mv"] + (1 - p) * V[k]
            t["Mvv"] = p * t["Mvv"] + (1 - p) * np.outer(V[k], V[k])
        t["m"] = m
        p = m / (m + 1.0)
        t["SigmaLm"] = o.chol_fullrank((2.38 ** 2 / 116 / p) * (t["Mvv"] - np.outer(t["Mv"], t["Mv"])))
        assert t["SigmaLm"] is not None
        return t

    for fresh in (False, True):
        z = rng.standard_normal((iters, 116, Cn)) * (1.0 if fresh else 0.3)
        u = rng.random((iters, Cn))
        um = rng.random((iters, Cn))
        u[1] = 1e-300                      # force an accept
        u[3] = 1.0                         # force a reject
        um[2] = 0.01                       # below beta: the initial proposal even past 2d
        tunes = [o.amm_new_tune(x0, np.eye(116) * sd) if fresh else history_tune(100 + k, 300) for k in range(Cn)]
        am = np.array([-1 if fresh else t["m"] for t in tunes], dtype=np.int64)
        aMv = np.stack([t["Mv"] for t in tunes])
        aMvv = np.stack([t["Mvv"] for t in tunes])
        aLm = np.stack([t["SigmaLm"] for t in tunes])
        state = L.fcol(np.tile(x0, (Cn, 1)))
        lf = np.full(Cn, np.nan)
        out = np.empty((iters, 116, Cn), order="F")
        nacc = np.zeros(Cn, dtype=np.int64)
        pr = c.prior()
        op = L.opts(1e-9, 1e-9)
        D = np.asfortranarray(gold["data"])
        sig = np.full(4, 0.1)
        pat = np.arange(Cn, dtype=np.int32)
        tune = L.AmmTune(0.05, 2.38, am.ctypes.data_as(L.c_int64_p), L.dptr(aMv), L.dptr(aMvv), L.dptr(aLm))
        L.check(lib.gync_mcmc_run_amm(L.dptr(state), L.dptr(lf), Cn, None, sd, L.dptr(D), 4, L.dptr(sig),
                                      pat.ctypes.data_as(L.c_int32_p), C.byref(pr), C.byref(op), iters, 1, 0, 0,
                                      L.dptr(np.asfortranarray(z)), L.dptr(np.asfortranarray(u)), L.dptr(np.asfortranarray(um)),
                                      L.dptr(out), nacc.ctypes.data_as(L.c_int64_p), C.byref(tune)))
        for cidx in range(Cn):
            xs, t = x0.copy(), tunes[cidx]
            cur = logf(xs, cidx)
            acc = 0
            for it in range(iters):
                xs, cur, a = o.amm_step(xs, cur, t, z[it, :, cidx], u[it, cidx], um[it, cidx], lambda v: logf(v, cidx))
                acc += a
                # the re-factorised SigmaLm amplifies rounding by the conditioning of Sigma (two different Cholesky orders)
                assert relerr(out[it, :, cidx], np.exp(xs)) < 1e-9, (fresh, cidx, it)
            assert acc == nacc[cidx] and acc >= 1
            assert am[cidx] == t["m"]
            # moments inherit the state differences: rounding only (fma contraction) without adapted proposals, the
            # factor-conditioning bound above with them
            tolm = dict(rtol=1e-12, atol=1e-13) if fresh else dict(rtol=1e-9, atol=1e-10)
            assert np.allclose(aMv[cidx], t["Mv"], **tolm)
            assert np.allclose(aMvv[cidx], t["Mvv"], **tolm)
            if not fresh:
                # the factor amplifies rounding by the conditioning of Sigma; compare the covariance it represents
                assert np.allclose(aLm[cidx] @ aLm[cidx].T, t["SigmaLm"] @ t["SigmaLm"].T, rtol=1e-7, atol=1e-12)
                assert np.allclose(aLm[cidx], np.tril(aLm[cidx]))
            else:
                assert np.array_equal(aLm[cidx], np.eye(116) * sd)   # rank-deficient so far: SigmaLm is still SigmaL


def test_sample_api_adaptive(gpu):
    """sample(Config(adapt=true), n): reproducible, resumable, independent of the speculation depth; adaptation
    kicks in after 2d = 232 iterations (propadapt becomes a full-rank covariance)."""
    c = gpu.Config(gpu.Lausanne(1), adapt=True)
    s = gpu.sample(c, 150, seed=5)
    assert s.samples.shape == (150, 116) and s.variate.m == 150
    assert np.allclose(gpu.api.propinit(s), c.propvar, rtol=1e-14, atol=0)
    s = gpu.sample_bang(s, 150)
    assert s.samples.shape == (300, 116) and s.variate.m == 300
    s2 = gpu.sample(c, 300, seed=5)
    assert np.array_equal(s.samples, s2.samples)                       # resumable: 150 + 150 == 300
    os.environ["GYNC_MCMC_SPEC"] = "3"
    try:
        s3 = gpu.sample(c, 300, seed=5)
    finally:
        del os.environ["GYNC_MCMC_SPEC"]
    assert np.array_equal(s3.samples, s.samples)                       # speculation is exact under adaptation too
    P = gpu.api.propadapt(s)
    nuniq = len(np.unique(s.samples[:, 0]))
    assert nuniq >= 2 and np.all(np.linalg.eigvalsh(P) > 0)
    if nuniq <= 116:                                                   # fewer than d+1 distinct states: rank-deficient estimate
        assert np.allclose(P, c.propvar)                               # -> the adapted component is still the initial proposal
    assert np.allclose(s.variate.Mv, np.log(np.vstack([gpu.api.init(c), s.samples])).mean(0), rtol=1e-10)


def test_checkpoint_resume(gpu, tmp_path):
    """save / load / batch (utils.jl:48-49, batch.jl:26-46): a chain continued from its checkpoint equals the
    uninterrupted one, with and without adaptation."""
    for adapt in (False, True):
        c = gpu.Config(gpu.Lausanne(2), thin=2, adapt=adapt)
        path = str(tmp_path / c.filename())
        s = gpu.api.batch(c, [10, 30], path, seed=4)
        assert s.samples.shape == (15, 116) and os.path.isfile(path)
        s = gpu.api.batch(c, [30, 50], path)                       # continues the file: 20 more iterations
        assert s.samples.shape == (25, 116)
        ref = gpu.sample(c, 50, seed=4)
        assert np.array_equal(s.samples, ref.samples) and s.variate.naccept == ref.variate.naccept
        l = gpu.api.load(path)
        assert l.config.adapt == adapt and l.config.patient.id == "l2" and l.variate.m == (50 if adapt else -1)


def test_sample_api_shapes_and_reproducibility(gpu):
    c = gpu.Config(gpu.Lausanne(1), thin=2)
    s = gpu.sample(c, 20, seed=3)
    assert s.samples.shape == (10, 116)
    s = gpu.sample_bang(s, 20)
    assert s.samples.shape == (20, 116)                                         # test/gync.jl:14
    s2 = gpu.sample_bang(gpu.sample(c, 20, seed=3), 20)
    assert np.array_equal(s.samples, s2.samples)                                # counter-based RNG: resumable, reproducible
    assert 0 < s.variate.naccept <= 40 and np.all(s.samples > 0)
    # speculative evaluation is EXACT: depth 1 (the sequential algorithm) and the default depth give the same chain
    os.environ["GYNC_MCMC_SPEC"] = "1"
    try:
        s1 = gpu.sample_bang(gpu.sample(c, 20, seed=3), 20)
    finally:
        del os.environ["GYNC_MCMC_SPEC"]
    assert np.array_equal(s1.samples, s.samples) and s1.variate.naccept == s.variate.naccept
    ss = gpu.sample_chains([gpu.new_sampling(gpu.Config(gpu.Lausanne(i)), seed=1) for i in (1, 2, 3)], 6)
    assert all(x.samples.shape == (6, 116) for x in ss)
    assert not np.array_equal(ss[0].samples, ss[1].samples)
    w = gpu.WeightedChain([s, s])                                               # test/gync.jl:27-29
    assert w.likelihoods.shape[1] == 2 and abs(w.weights.sum() - 1) < 1e-12
    assert gpu.sample(w, 10).shape == (10, 116)
    gpu.emiteration_bang(w)


def test_persistent_kernel_equals_round_loop(gpu):
    """The persistent Metropolis kernel (gync_mcmc_persist.cuh: chains advance out of step, rounds re-proposed on the
    device) produces bit-identical chains to the round-1 host loop of separate propose / solve / accept launches
    (GYNC_MCMC_ROUNDS=1), for one chain and for several chains with different patients, thinned and not."""
    for thin, iters, pats in ((1, 40, (1,)), (3, 45, (1, 2, 3, 4, 5))):
        runs = []
        for rounds in (False, True):
            if rounds:
                os.environ["GYNC_MCMC_ROUNDS"] = "1"
            try:
                ss = gpu.sample_chains([gpu.new_sampling(gpu.Config(gpu.Lausanne(i), thin=thin), seed=21) for i in pats], iters)
            finally:
                os.environ.pop("GYNC_MCMC_ROUNDS", None)
            runs.append(ss)
        for a, b in zip(*runs):
            assert a.samples.shape == (iters // thin, 116)
            assert np.array_equal(a.samples, b.samples) and a.variate.naccept == b.variate.naccept
            assert a.variate.logf == b.variate.logf and np.array_equal(a.variate.state, b.variate.state)
        assert sum(x.variate.naccept for x in runs[0]) > 0
    solves, steps, k = C.c_uint64(), C.c_uint64(), C.c_int32()
    gpu.lib.load().gync_mcmc_last_stats(C.byref(solves), C.byref(steps), C.byref(k))
    assert k.value >= 1        # the last call above was a round-loop run: the counters belong to the persistent run before it


def test_persistent_kernel_high_acceptance_and_edge_cases(gpu):
    """The cancellation machinery under stress and the corners: a small proposal (acceptance ~50 %: almost every window is
    cancelled and re-issued, generations turn over every other iteration) for one chain with a deep window and for many
    chains with a shallow one, bit-identical to the round loop; thin > iters (no output rows); an all-missing patient
    table (every proposal scores -Inf: the chain never moves, model.jl:161); zero iterations."""
    small = gpu.api.uniformpropvar(0.0007)
    nan_patient = gpu.Patient(np.full((31, 4), np.nan), "none")
    for pats, iters, thin in (((1,), 60, 1), (tuple(range(1, 25)), 30, 2)):
        runs = []
        for rounds in (False, True):
            if rounds:
                os.environ["GYNC_MCMC_ROUNDS"] = "1"
            try:
                ss = gpu.sample_chains([gpu.new_sampling(gpu.Config(gpu.Lausanne(i), propvar=small, thin=thin), seed=33) for i in pats], iters)
            finally:
                os.environ.pop("GYNC_MCMC_ROUNDS", None)
            runs.append(ss)
        for a, b in zip(*runs):
            assert np.array_equal(a.samples, b.samples) and a.variate.naccept == b.variate.naccept and a.variate.logf == b.variate.logf
        acc = np.mean([x.variate.naccept for x in runs[0]]) / iters
        assert 0.15 < acc < 0.95, acc
    s = gpu.sample(gpu.Config(gpu.Lausanne(1), thin=50), 20, seed=2)
    assert s.samples.shape == (0, 116) and s.variate.iteration == 20
    s = gpu.sample_bang(s, 30)
    assert s.samples.shape == (0, 116) and s.variate.iteration == 50          # floor(30/50) rows per call, like the reference
    dead = gpu.sample(gpu.Config(nan_patient), 15, seed=2)
    assert dead.variate.naccept == 0 and dead.variate.logf == -np.inf
    assert np.allclose(dead.samples, np.tile(gpu.api.init(dead.config), (15, 1)), rtol=1e-15)
    z = gpu.sample(gpu.Config(), 0, seed=2)
    assert z.samples.shape == (0, 116)


def test_chains_sharded_over_devices_in_one_process(gpu):
    """gync_init_multi: the library gives every device a contiguous block of the chains and its own persistent kernel; the
    chains are bit-identical to the same chains on one device (their Philox streams belong to them).  Needs >= 2 GPUs."""
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    lib, L = gpu.lib.load(), gpu.lib
    mk = lambda: [gpu.new_sampling(gpu.Config(gpu.Lausanne(1 + i % 5), thin=2), seed=8) for i in range(3 * nd + 1)]
    one = gpu.sample_chains(mk(), 20)
    L.check(lib.gync_init_multi(nd, None))
    try:
        many = gpu.sample_chains(mk(), 20)
    finally:
        L.check(lib.gync_init(0))
    for a, b in zip(one, many):
        assert np.array_equal(a.samples, b.samples) and a.variate.naccept == b.variate.naccept and a.variate.logf == b.variate.logf


def test_chain_keeps_its_random_stream_when_resumed_alone(gpu, tmp_path):
    """A chain advanced inside a group and later resumed ALONE (sample!, batch, load) continues the stream it would have
    drawn from inside the group: the Philox stream id is the variate's chain_id, persisted by save/load, not the
    position in the call.  Chains of one group started with the same seed do not share their noise."""
    mk = lambda: [gpu.new_sampling(gpu.Config(gpu.Lausanne(i)), seed=9) for i in (1, 1, 2)]
    ref = gpu.sample_chains(mk(), 24)
    assert [x.variate.chain_id for x in ref] == [0, 1, 2]
    assert not np.array_equal(ref[0].samples, ref[1].samples)            # same patient, same seed, different stream
    ss = gpu.sample_chains(mk(), 12)
    path = str(tmp_path / "chain2.jld")
    gpu.api.save(path, ss[2])
    alone = gpu.sample_bang(gpu.api.load(path), 12)                       # position 0 of its own call, stream 2
    assert alone.variate.chain_id == 2
    assert np.array_equal(alone.samples, ref[2].samples) and alone.variate.naccept == ref[2].variate.naccept
    pair = gpu.sample_chains([ss[1], ss[0]], 12)                          # regrouped in another order
    assert np.array_equal(pair[0].samples, ref[1].samples) and np.array_equal(pair[1].samples, ref[0].samples)
    # chains created without a seed get the process seed and distinct stream ids
    a, b = gpu.new_sampling(gpu.Config()), gpu.new_sampling(gpu.Config())
    assert a.variate.seed == b.variate.seed and a.variate.chain_id != b.variate.chain_id


def test_mcmc_statistics_vs_independent_cpu_metropolis(gpu, gold):
    """the chains are not only step-exact with injected draws, they are the right Markov
    chain with their OWN randomness.  512 device chains (Philox) and 32 independent CPU chains (numpy PCG64, the oracle's
    metropolis_step, log-target through the C oracle at 1e-8) start from the reference point on the real target
    (Lausanne 1, sigma_rho = 0.1) and run 150 iterations; acceptance counts, the squared displacement and every
    coordinate's mean of the final state must agree within Monte-Carlo error (independent chains: MCSE = s / sqrt(n))."""
    import c_oracle as co
    import gync_oracle as o
    means, stds = gold["gmm_means"], gold["gmm_stds"]
    data = o.patients()[0][0]
    iters, n_gpu, n_cpu = 150, 512, 32
    cfg = gpu.Config(gpu.Patient(data, "l1"), priory0=gpu.api.GaussianMixture(means, stds))
    ss = gpu.sample_chains([gpu.new_sampling(cfg, seed=123) for _ in range(n_gpu)], iters)
    x0 = np.log(o.init_x())
    XG = np.log(np.array([s_.samples[-1] for s_ in ss]))
    AG = np.array([s_.variate.naccept for s_ in ss], dtype=float)

    rng = np.random.Generator(np.random.PCG64(2024))
    sd = np.sqrt(np.log(1.01))
    SL = np.eye(116) * sd
    x = np.tile(x0, (n_cpu, 1))

    def logf_batch(XL):
        X = np.exp(XL)
        lp = np.array([o.logprior(v, means, stds) for v in X])
        out = np.full(len(X), -np.inf)
        ok = np.isfinite(lp)
        if ok.any():
            out[ok] = lp[ok] + co.llh_batch(X[ok], [data], [0.1], 1, 1e-8, 1e-8)[:, 0] + XL[ok].sum(1)
        return out
    cur = logf_batch(x)
    AC = np.zeros(n_cpu)
    for _ in range(iters):
        z, u = rng.standard_normal((n_cpu, 116)), rng.random(n_cpu)
        lpn = logf_batch(x + z @ SL.T)
        for c_ in range(n_cpu):
            x[c_], cur[c_], a = o.metropolis_step(x[c_], cur[c_], SL, z[c_], u[c_], lambda v, c_=c_: lpn[c_])
            AC[c_] += a

    def close(g_, c_, what, k=4.5):
        se = np.sqrt(g_.var(ddof=1) / len(g_) + c_.var(ddof=1) / len(c_))
        assert abs(g_.mean() - c_.mean()) <= k * se, (what, g_.mean(), c_.mean(), se)
    assert 0.02 < AG.mean() / iters < 0.5
    close(AG, AC, "acceptance count")
    close(((XG - x0) ** 2).sum(1), ((x - x0) ** 2).sum(1), "squared displacement")
    se = np.sqrt(XG.var(0, ddof=1) / n_gpu + x.var(0, ddof=1) / n_cpu)
    zscore = np.abs(XG.mean(0) - x.mean(0)) / se
    assert zscore.max() < 5.0, (int(zscore.argmax()), float(zscore.max()))          # 116 coordinates: max |z| ~ 3 expected
    assert np.mean(zscore ** 2) < 1.6                                               # and the z-scores are N(0,1)-sized as a set
    vr = XG.var(0, ddof=1).sum() / x.var(0, ddof=1).sum()
    assert 0.6 < vr < 1.6, vr                                                       # total variance of the final states


def test_em_resident_matrix_handle(gpu):
    """emiteration!(c::WeightedChain) driven one iteration per call (em.jl:7): the likelihood matrix is uploaded once and stays
    in HBM (gync_em_create / gync_em_iterate_handle); the iterates equal the oracle's and the one-shot host-array call."""
    import gync_oracle as o
    rng = np.random.default_rng(4)
    K, M = 20_011, 40
    Lm = rng.random((K, M))
    Lm /= Lm.sum(0, keepdims=True)
    wc = gpu.WeightedChain(rng.random((K, 116)), Lm, np.full(K, 1.0 / K), np.ones(K))
    ref = np.full(K, 1.0 / K)
    l0 = gpu.lib.load().gync_launch_count()
    for _ in range(7):
        gpu.emiteration_bang(wc)
        ref = o.emiteration(ref, Lm)
    assert relerr(wc.weights, ref) < RTOL_EM and abs(wc.weights.sum() - 1) < 1e-12
    assert wc._em is not None and gpu.lib.load().gync_launch_count() - l0 == 7
    w2 = np.full(K, 1.0 / K)
    gpu.emiteration_bang(w2, Lm, niter=7)
    assert relerr(w2, wc.weights) < 1e-13
    h0 = wc._em.value
    wc.likelihoods = np.asfortranarray(Lm[:, ::-1].copy())          # rebound: the handle follows
    gpu.emiteration_bang(wc)
    assert wc._em.value != h0 or wc._em_key[0] == wc.likelihoods.ctypes.data
    wc.forget_device_matrix()
    assert wc._em is None


def test_wc_normalize_vs_oracle(gpu):
    import gync_oracle as o
    lib, L = gpu.lib.load(), gpu.lib
    rng = np.random.default_rng(9)
    for K, M in ((5, 1), (1000, 7), (30011, 40)):
        LL = np.asfortranarray(-rng.random((K, M)) * 800)
        LL[rng.integers(K), rng.integers(M)] = -np.inf
        lp = -rng.random(K) * 50
        counts = rng.integers(1, 9, K).astype(np.int64)
        Lm, w, d = np.empty((K, M), order="F"), np.empty(K), np.empty(K)
        L.check(lib.gync_wc_normalize(L.dptr(LL), L.dptr(lp), counts.ctypes.data_as(L.c_int64_p), K, M, L.dptr(Lm),
                                      L.dptr(w), L.dptr(d)))
        l0, w0, d0 = o.weightedchain_normalize(lp, LL, counts)
        nz = l0 > 0
        assert np.array_equal(Lm == 0, l0 == 0) and relerr(Lm[nz], l0[nz]) < 1e-12
        assert relerr(w, w0) < 1e-14 and relerr(d, d0) < 1e-12


def test_wc_normalize_sharded_kernels(gpu):
    """SURVEY §8e row K5 on one GPU: two row shards driven through dist.wc_normalize_sharded with the collectives
    emulated by combining the shards' buffers; equals the unsharded oracle."""
    import torch
    import gync_oracle as o
    from gync_b200.dist import DeviceWC
    rng = np.random.default_rng(19)
    K, M = 20011, 40
    LL = -rng.random((K, M)) * 600
    LL[rng.integers(K), rng.integers(M)] = -np.inf
    lp = -rng.random(K) * 50
    counts = rng.integers(1, 9, K).astype(np.int64)
    dev = torch.device("cuda", 0)
    cut = 7003
    shards = []
    for lo, hi in ((0, cut), (cut, K)):
        shards.append(DeviceWC(torch.from_numpy(np.ascontiguousarray(LL[lo:hi].T)).to(dev),
                               torch.from_numpy(lp[lo:hi].copy()).to(dev), torch.from_numpy(counts[lo:hi].copy()).to(dev)))
    # lock-step emulation of the four collectives
    st = [s.colstats() for s in shards]
    torch.cuda.synchronize()
    g = torch.maximum(st[0][:M + 1], st[1][:M + 1])
    tot = st[0][M + 1:] + st[1][M + 1:]
    for x in st:
        x[:M + 1] = g
        x[M + 1:] = tot
    cs = [s.expsum(x) for s, x in zip(shards, st)]
    torch.cuda.synchronize()
    gs = cs[0] + cs[1]
    for x in st:
        x[:M] = gs
    tt = [s.rows(x) for s, x in zip(shards, st)]
    torch.cuda.synchronize()
    gt = tt[0] + tt[1]
    for s in shards:
        s.scale(gt)
    torch.cuda.synchronize()
    l0, w0, d0 = o.weightedchain_normalize(lp, LL, counts)
    Lg = np.vstack([s.LL.cpu().numpy().T for s in shards])
    wg = np.concatenate([s.weights.cpu().numpy() for s in shards])
    dg = np.concatenate([s.density.cpu().numpy() for s in shards])
    nz = l0 > 0
    assert np.array_equal(Lg == 0, l0 == 0) and relerr(Lg[nz], l0[nz]) < 1e-12
    assert relerr(wg, w0) < 1e-14 and relerr(dg, d0) < 1e-12


def test_learned_queue_order_is_result_neutral(gpu):
    """csrc/gync_sched.cuh: from the second batch on the sample queue is ordered by a cost model refitted on the device;
    log-likelihoods, step counts and statuses must be bit-identical to the unordered run (GYNC_NO_SCHED=1)."""
    X = near_set(40_000, 77)
    X[123, 5] = np.nan                      # a sample that fails at load time is scheduled like any other
    c = gpu.Config()
    runs = {}
    for name, env in (("first", None), ("ordered", None), ("unordered", "1")):
        if env:
            os.environ["GYNC_NO_SCHED"] = env
        try:
            runs[name] = gpu.loglik_matrix(X, [c.data], [0.1], want_steps=True)
        finally:
            os.environ.pop("GYNC_NO_SCHED", None)
    for name in ("first", "ordered"):
        for a, b in zip(runs[name], runs["unordered"]):
            assert np.array_equal(a, b, equal_nan=True), name
    assert runs["ordered"][1][123] != 0 and runs["ordered"][0][123, 0] == -np.inf


def test_collapse_duplicates_device(gpu):
    """Head of WeightedChain(samplings) (weightedchain.jl:31-46) on the device vs the oracle / host mirror."""
    import gync_oracle as o
    rng = np.random.default_rng(29)
    base = rng.random((4000, 116)) + 0.1
    reps = rng.integers(1, 6, 4000)                       # runs of repeated rows, as a Metropolis chain produces them
    S = np.repeat(base, reps, axis=0)
    cut = S.shape[0] // 3
    c = gpu.Config()
    ss = [gpu.Sampling(S[:cut], c, None), gpu.Sampling(S[cut:], c, None)]     # a run may continue across the chain boundary
    rows, counts = gpu.api.collapse_duplicates_device(ss)
    r0, c0 = o.collapse_duplicates([S[:cut], S[cut:]])
    assert np.array_equal(rows, r0) and np.array_equal(counts, c0) and counts.sum() == S.shape[0]
    r1, c1 = gpu.api.collapse_duplicates(ss)
    assert np.array_equal(rows, r1) and np.array_equal(counts, c1)
    rows, counts = gpu.api.collapse_duplicates_device(ss, burnin=7)
    r0, c0 = o.collapse_duplicates([S[:cut], S[cut:]], burnin=7)
    assert np.array_equal(rows, r0) and np.array_equal(counts, c0)
    one = [gpu.Sampling(S[:1], c, None)]
    rows, counts = gpu.api.collapse_duplicates_device(one)
    assert rows.shape == (1, 116) and counts.tolist() == [1]
