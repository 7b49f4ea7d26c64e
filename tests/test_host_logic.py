"""CPU suite, part 2: generated code + solver core compiled for the host (tests/host_shim), the C-ABI
library's symbol table, and the host-side mirror of the reference API.  No compute call on a GPU."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import c_oracle as co
import gync_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
PARITY_RTOL, PARITY_ATOL = 7e-10, 1e-11     # library defaults (include/gync_b200.h, api.DEFAULT_RTOL/ATOL)
dp = C.POINTER(C.c_double)
P = lambda a: a.ctypes.data_as(dp)  # noqa: E731


def relerr(a, b):
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))


def test_generated_header_is_reproducible(tmp_path):
    """csrc/gen/gync_model_gen.h is the generator's output, byte for byte (nobody edits it by hand; the knobs of
    codegen/generate.py — load distances of the straight-line code — are at their committed defaults)."""
    import subprocess, sys
    pytest.importorskip("sympy")
    out = tmp_path / "gen.h"
    env = {k: v for k, v in os.environ.items() if not k.startswith("GYNC_GEN_")}
    env["GYNC_GEN_OUT"] = str(out)
    subprocess.run([sys.executable, os.path.join(ROOT, "gync.jl_b200", "codegen", "generate.py")], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    committed = open(os.path.join(ROOT, "gync.jl_b200", "csrc", "gen", "gync_model_gen.h")).read()
    assert out.read_text() == committed


def test_generated_rhs_vs_reference_text(host_shim):
    g = np.load(os.path.join(G, "rhs_golden.npz"))
    for y, p, f in zip(g["Y"], g["P"], g["F"]):
        out = np.empty(33)
        host_shim.hs_rhs(P(y), P(p), P(out))
        # reciprocal-threshold form: a few ulp on top of the cancellation already in f
        assert np.max(np.abs(out - f) / (np.abs(f) + 1e-12 * np.abs(f).max())) < 1e-11


def test_generated_jacobian_and_sparse_lu(host_shim):
    g = np.load(os.path.join(G, "rhs_golden.npz"))
    rng = np.random.default_rng(0)
    nq, nlu = C.c_int(), C.c_int()
    assert host_shim.hs_dims(C.byref(nq), C.byref(nlu)) == 33 and nlu.value == 165
    for y, p in zip(g["Y"][:24], g["P"][:24]):
        Jref = np.empty((33, 33))
        co.lib().orc_jac_exact(P(y), P(p), P(Jref))           # oracle: complex-step Jacobian
        for fac in (1e5, 400.0, 10.0, 0.5):
            f = np.empty(33)
            Wd = np.empty((33, 33))
            host_shim.hs_rhs_w(P(y), P(p), C.c_double(fac), P(f), P(Wd))
            Wref = fac * np.eye(33) - Jref
            assert np.max(np.abs(Wd - Wref) / (np.abs(Wref) + 1e-9 * np.abs(Wref).max())) < 1e-12
            b = rng.standard_normal(33) * (np.abs(f) + 1e-3)
            x = b.copy()
            host_shim.hs_lusolve(P(y), P(p), C.c_double(fac), P(x))     # no-pivot sparse LU (static Markowitz order)
            xr = np.linalg.solve(Wref, b)
            assert np.max(np.abs(x - xr)) / np.max(np.abs(xr)) < 1e-10


def host_traj(host_shim, x, rtol, atol):
    tr = np.empty((62, 4))
    na, nr = C.c_int(), C.c_int()
    st = host_shim.hs_meas_traj(P(np.ascontiguousarray(x)), C.c_double(rtol), C.c_double(atol), 200000, P(tr),
                                C.byref(na), C.byref(nr))
    return st, tr, na.value, nr.value


def test_rodas4_solver_core_vs_converged_golden(host_shim):
    """The device solver's source, compiled for the CPU, against the converged oracle values:
    trajectories within 1e-6, log-likelihood within 1e-8 relative (north-star tolerances)."""
    g = np.load(os.path.join(G, "solution_golden.npz"))
    for i in (0, 1, 2, 9, 34):
        st, tr, na, nr = host_traj(host_shim, g["X"][i], PARITY_RTOL, PARITY_ATOL)
        assert st == 0 and 300 < na < 20000
        assert relerr(tr, g["traj"][i]) < 1e-6
        for m in range(4):
            l = o.llh_from_traj(tr, g["data"][:, :, m], 0.1)
            assert abs(l - g["LL"][i, m]) <= 1e-8 * abs(g["LL"][i, m]), (i, m, l, g["LL"][i, m])


def test_solver_core_at_the_reference_tolerance(host_shim):
    """A caller may keep GynC.jl's own tolerance (reltol = abstol = 1e-4, gyncycle.jl:8) instead of the library default:
    the solver core then lands in the accuracy class of the reference-tolerance BDF (oracle, frozen in the goldens) — both
    are 1e-4 .. 1e-2 relative away from the converged log-likelihood — in ~175 step attempts instead of ~4 100."""
    g = np.load(os.path.join(G, "solution_golden.npz"))
    ours, bdf, steps = [], [], []
    for i in range(8):
        st, tr, na, nr = host_traj(host_shim, g["X"][i], 1e-4, 1e-4)
        assert st == 0
        l = o.llh_from_traj(tr, g["data"][:, :, 0], 0.1)
        ours.append(abs(l - g["LL"][i, 0]) / abs(g["LL"][i, 0]))
        bdf.append(abs(g["LL_bdf1e4"][i, 0] - g["LL"][i, 0]) / abs(g["LL"][i, 0]))
        steps.append(na + nr)
    assert max(ours) < 1e-2 and max(ours) < 2 * max(bdf), (ours, bdf)
    assert 100 < np.mean(steps) < 400, steps


def test_solver_core_on_the_parity_tail_fixtures(host_shim):
    """CPU build of the device solver at the library default tolerance against the frozen near-set log-likelihoods
    (tests/golden/nearset_llh_golden.npz): a slice of both populations plus the two samples that were above the 1e-8
    bound in round 1 (RODAS4 at 1e-8/1e-10: 4.6e-8 and 1.1e-8), and a slice of the wide set incl. its failure set.
    The -m gpu suite runs all 30 720 + 512 cases."""
    sys.path.insert(0, G)
    import make_nearset_golden as mg
    g = np.load(os.path.join(G, "nearset_llh_golden.npz"))
    la = o.patients()[0]
    worst = 0.0
    for name, idx in (("A", list(range(40)) + [187, 963, 740]), ("B", list(range(40)) + [2730, 2929]), ("C", list(range(40)))):
        seed, n, pats = mg.SETS[name]
        X = mg.near_set(seed, n)
        for i in idx:
            st, tr, _, _ = host_traj(host_shim, X[i], PARITY_RTOL, PARITY_ATOL)
            assert st == 0
            for m, pi in enumerate(pats):
                l = o.llh_from_traj(tr, la[pi], 0.1)
                worst = max(worst, abs(l - g["LL_" + name][i, m]) / abs(g["LL_" + name][i, m]))
    assert worst <= 1e-8, worst
    Xw = mg.wide_set(int(g["seed_W"]), 512)
    for i in range(64):
        st, tr, _, _ = host_traj(host_shim, Xw[i], PARITY_RTOL, PARITY_ATOL)
        assert (st == 0) == bool(np.isfinite(g["LL_W"][i])), i
        if st == 0:
            l = o.llh_from_traj(tr, la[0], 0.1)
            assert abs(l - g["LL_W"][i]) <= 1e-8 * abs(g["LL_W"][i])


def test_rodas4_failure_status(host_shim):
    g = np.load(os.path.join(G, "solution_golden.npz"))
    bad = np.flatnonzero(~np.isfinite(g["LL"][:, 0]))[0]
    st, tr, _, _ = host_traj(host_shim, g["X"][bad], PARITY_RTOL, PARITY_ATOL)
    assert st != 0 and np.isnan(tr).all()
    x = g["X"][0].copy()
    x[115] = np.nan
    st, tr, _, _ = host_traj(host_shim, x, 1e-6, 1e-6)
    assert st == 4 and np.isnan(tr).all()


def _tableau():
    """RD_* macros of csrc/gync_solver.cuh -> (gamma, A, C) of the 6-stage implementation form."""
    src = open(os.path.join(ROOT, "gync.jl_b200", "csrc", "gync_solver.cuh")).read()
    co_ = {m.group(1): float(m.group(2).strip("()")) for m in re.finditer(r"#define RD_([AGC]\d*)\s+(\(?-?[0-9.e+-]+\)?)", src)}
    s = 6
    A, Cm = np.zeros((s, s)), np.zeros((s, s))
    for k, v in co_.items():
        if k[0] == "A":
            A[int(k[1]) - 1, int(k[2]) - 1] = v
        elif k[0] == "C":
            Cm[int(k[1]) - 1, int(k[2]) - 1] = v
    A[5, :4] = A[4, :4]
    A[5, 4] = 1.0          # stage 6 starts from the embedded solution (gync_rodas4_step, RD_TA last row)
    return co_["G"], A, Cm


def test_rosenbrock_tableau_order_conditions():
    """The 6-stage alternative's coefficient table (RODAS4P, -DGYNC_ROS_ORDER=4 builds) is pinned by the Rosenbrock order conditions themselves
    (Hairer & Wanner II, Table IV.7.1): all 8 conditions up to order 4 for the propagated weights, the 4 conditions up
    to order 3 for the embedded weights, stiff accuracy (alpha_6 = 1, last stage weight 1) — a wrong digit in any
    coefficient breaks them at the 1e-6 level, they hold to 1e-13."""
    g, A, Cm = _tableau()
    Gam = np.linalg.inv(np.eye(6) / g - Cm)                  # Gamma: gamma on the diagonal, gamma_ij below
    assert np.allclose(np.diag(Gam), g, atol=1e-14) and np.allclose(np.triu(Gam, 1), 0, atol=1e-14)
    al = A @ Gam
    Gs = Gam - np.diag(np.diag(Gam))
    be = al + Gs
    ai, bi = al.sum(1), be.sum(1)
    assert abs(ai[5] - 1.0) < 1e-13 and abs(ai[4] - 1.0) < 1e-13
    m = np.array([A[4, 0], A[4, 1], A[4, 2], A[4, 3], 1.0, 1.0])

    def cond(b):
        return np.array([b.sum() - 1, b @ bi - (0.5 - g), b @ ai ** 2 - 1 / 3, b @ be @ bi - (1 / 6 - g + g * g),
                         b @ ai ** 3 - 0.25, b @ (ai * (al @ bi)) - (1 / 8 - g / 3), b @ be @ ai ** 2 - (1 / 12 - g / 3),
                         b @ be @ be @ bi - (1 / 24 - g / 2 + 1.5 * g * g - g ** 3)])
    assert np.max(np.abs(cond(m @ Gam))) < 1e-13
    mh = m.copy()
    mh[5] = 0.0
    ch = cond(mh @ Gam)
    assert np.max(np.abs(ch[:4])) < 1e-13 and np.max(np.abs(ch[4:])) > 1e-3     # embedded: order 3, not 4


def _tableau5():
    """R5_* macros of csrc/gync_solver.cuh -> (gamma, A, C, weights, embedded weights) of the 8-stage implementation form
    of Rodas5P, with the structure the step function relies on: stage 7 = stage-6 point + k6, stage 8 = stage-7 point + k7,
    new solution = stage-8 point + k8, error estimate = k8."""
    src = open(os.path.join(ROOT, "gync.jl_b200", "csrc", "gync_solver.cuh")).read()
    co_ = {m.group(1): float(m.group(2).strip("()")) for m in re.finditer(r"#define R5_([AGC]\d*)\s+(\(?-?[0-9.e+-]+\)?)", src)}
    A, Cm = np.zeros((8, 8)), np.zeros((8, 8))
    for k, v in co_.items():
        if k[0] == "A":
            A[int(k[1]) - 1, int(k[2]) - 1] = v
        elif k[0] == "C":
            Cm[int(k[1]) - 1, int(k[2]) - 1] = v
    assert np.count_nonzero(A) == 15 and np.count_nonzero(Cm) == 28
    A[6, :5] = A[5, :5]
    A[6, 5] = 1.0
    A[7, :6] = A[6, :6]
    A[7, 6] = 1.0
    m = A[7].copy()
    m[7] = 1.0
    return co_["G"], A, Cm, m, A[7].copy(), co_


def test_rodas5p_tableau_order_conditions():
    """The default method's table (Rodas5P, written from memory of the publication) is pinned by the Rosenbrock order
    conditions themselves: tools/rodas5/order_conditions.py generates all 17 rooted trees up to order 5; the propagated
    weights satisfy every one of them and the embedded weights the 8 up to order 4 — and NOT order 5 — at round-off.  A
    wrong digit in any of the 44 coefficients breaks them at the 1e-6 level.  Also: stiff accuracy, and the identities
    behind the five-vector storage (k6 absorbed into the stored k4, k5; gync_solver.cuh)."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "rodas5"))
    import order_conditions as oc
    g, A, Cm, m, mh, co_ = _tableau5()
    res = oc.residuals(g, A, Cm, m, 5)
    assert len(res) == 17 and max(abs(r) for _, _, r in res) < 5e-13
    rh = oc.residuals(g, A, Cm, mh, 5)
    assert max(abs(r) for p_, _, r in rh if p_ <= 4) < 5e-13 and max(abs(r) for p_, _, r in rh if p_ == 5) > 1e-3
    al, be, b, Gam = oc.from_impl(g, A, Cm, m)
    assert np.allclose(np.diag(Gam), g, atol=1e-14) and np.allclose(np.triu(Gam, 1), 0, atol=1e-13)
    assert np.allclose(al.sum(1)[5:], 1.0, atol=1e-13)                     # stages 6..8 sit at t + h: stiffly accurate
    # a changed digit is caught
    A2 = A.copy()
    A2[3, 1] += 1e-7
    assert max(abs(r) for _, _, r in oc.residuals(g, A2, Cm, m, 5)) > 1e-9
    # the absorption of k6 into (k4, k5): a64 b + a65 a = 1, c84 b + c85 a = c86, and stage 7's corrected coefficient
    det = co_["C84"] * co_["A65"] - co_["C85"] * co_["A64"]
    beta, alpha = (co_["C86"] * co_["A65"] - co_["C85"]) / det, (co_["C84"] - co_["A64"] * co_["C86"]) / det
    assert abs(co_["A64"] * beta + co_["A65"] * alpha - 1) < 1e-13 and abs(co_["C84"] * beta + co_["C85"] * alpha - co_["C86"]) < 1e-12
    src = open(os.path.join(ROOT, "gync.jl_b200", "csrc", "gync_solver.cuh")).read()
    assert "#define R5_BETA  ((R5_C86 * R5_A65 - R5_C85) / R5_DET)" in src and "#define R5_ALPHA ((R5_C84 - R5_A64 * R5_C86) / R5_DET)" in src
    assert "R5_C76 - R5_C74 * R5_BETA - R5_C75 * R5_ALPHA" in src


def test_rodas5p_fixed_step_convergence(host_shim):
    """The step function as built (five stored stage vectors, k6 absorbed into k4 and k5) converges at least as fast as
    its order promises on the GynCycle ODE itself: with FIXED steps (hmax = h and a tolerance loose enough that the
    controller always proposes more) every halving of h shrinks the change of the log-likelihood at the reference point
    by more than 2^5, down to the round-off floor of ~1e-10 that tens of thousands of steps accumulate.  (The problem has no
    clean asymptotic regime — its LH surge makes the pre-asymptotic decay even faster — so this bounds the order from
    below; the order conditions above pin it exactly.)"""
    g_ = np.load(os.path.join(G, "solution_golden.npz"))
    la = o.patients()[0]
    x = np.ascontiguousarray(g_["X"][0])
    LLs = []
    for hmax in (0.032, 0.016, 0.008, 0.004):
        tr = np.empty((62, 4))
        na, nr = C.c_int(), C.c_int()
        st = host_shim.hs_meas_traj_hmax(P(x), C.c_double(1e-2), C.c_double(1e-2), C.c_double(hmax), 2000000, P(tr), C.byref(na), C.byref(nr))
        assert st == 0 and nr.value == 0 and na.value >= 58.0 / hmax
        LLs.append(o.llh_from_traj(tr, la[0], 0.1))
    d = np.abs(np.diff(LLs))
    assert d[0] / d[1] > 32 and d[1] / d[2] > 32, d
    assert abs(LLs[-1] - float(g_["LL"][0, 0])) < 1e-9 * abs(LLs[-1])          # and it converges to the golden value


def test_default_tolerances_agree_everywhere():
    hdr = open(os.path.join(ROOT, "include", "gync_b200.h")).read()
    rt = float(re.search(r"#define GYNC_DEFAULT_RTOL\s+(\S+)", hdr).group(1))
    at = float(re.search(r"#define GYNC_DEFAULT_ATOL\s+(\S+)", hdr).group(1))
    import gync_b200 as g
    assert (rt, at) == (PARITY_RTOL, PARITY_ATOL) == (g.api.DEFAULT_RTOL, g.api.DEFAULT_ATOL)
    jl = open(os.path.join(ROOT, "julia", "GynCB200.jl")).read()
    assert "reltol=7e-10, abstol=1e-11" in jl and "1e-8, 1e-10" not in jl


def test_rodas4_generic_grid_and_order(host_shim):
    g = np.load(os.path.join(G, "solution_golden.npz"))
    t = np.arange(0.0, 31.0)
    Y = np.empty((31, 33))
    p = o.allparms(o.REFPARMS)
    st = host_shim.hs_solve(P(o.REFY0), P(p), P(t), 31, C.c_double(1e-9), C.c_double(1e-11), 500000, P(Y))
    assert st == 0
    ref = co.gync(o.REFY0, p, t, mode=1, rtol=1e-12, atol=1e-12)
    big = np.abs(ref) > 1e-6 * np.abs(ref).max(0)
    assert np.max(np.abs(Y - ref)[big] / np.abs(ref)[big]) < 1e-6
    assert relerr(o.referencesolution(Y)[:, o.MEASUREDINDS], g["refsol"][:, o.MEASUREDINDS]) < 1e-7
    # repeated output time is served from the same state (CVODE semantics for a repeated tout)
    t2 = np.array([0.0, 1.0, 1.0, 2.0])
    Y2 = np.empty((4, 33))
    assert host_shim.hs_solve(P(o.REFY0), P(p), P(t2), 4, C.c_double(1e-8), C.c_double(1e-8), 100000, P(Y2)) == 0
    assert np.array_equal(Y2[1], Y2[2]) and np.array_equal(Y2[0], o.REFY0)


def test_philox_known_answers(host_shim):
    out = (C.c_uint32 * 4)()
    kat = [((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), 0x299f31d0a4093822,
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:      # Random123 kat_vectors, philox4x32-10
        host_shim.hs_philox(C.c_uint64(key), *ctr, out)
        assert tuple(out) == exp
    z = np.empty(116)
    u = C.c_double()
    zs = []
    for it in range(200):
        host_shim.hs_draw(C.c_uint64(42), C.c_uint64(3), C.c_uint64(it), P(z), C.byref(u))
        assert 0.0 < u.value < 1.0
        zs.append(z.copy())
    zs = np.array(zs)
    assert abs(zs.mean()) < 0.03 and abs(zs.std() - 1) < 0.03


# ---------------------------------------------------------------------------------------- C ABI
def test_cabi_library_exports_every_declared_symbol(built):
    from gync_b200 import lib as L
    hdr = open(os.path.join(ROOT, "include", "gync_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gync_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(L.EXPORTS)
    lib = L.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gync_abi_version() == 1
    dims = [C.c_int32() for _ in range(5)]
    lib.gync_dims(*[C.byref(d) for d in dims])
    assert [d.value for d in dims] == [33, 114, 116, 93, 165]
    assert lib.gync_flops_per_step() > 5000


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute call must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gync_b200 as g
    with pytest.raises(g.lib.GyncError, match="no CUDA device"):
        g.llh(g.Config(), g.api.init(g.Config()))
    with pytest.raises(g.lib.GyncError):
        g.emiteration(np.ones(4) / 4, np.ones((4, 2)))


def test_new_entry_points_validate_arguments_and_fail_without_a_device(built):
    """The round-2 entry points follow the ABI's error convention: bad arguments -> GYNC_ERR_ARG before any device work (so this
    runs without a GPU); good arguments without a CUDA device -> GYNC_ERR_CUDA with a message (no CPU fallback)."""
    import torch
    import gync_b200 as g
    L = g.lib
    lib = L.load()
    st, lf = np.zeros((2, 116), order="F"), np.full(2, np.nan)
    D, sig, pat = np.zeros((31, 4, 1), order="F"), np.array([0.1]), np.zeros(2, dtype=np.int32)
    op = L.opts(1e-6, 1e-8)
    ids = np.arange(2, dtype=np.uint64)
    idp = ids.ctypes.data_as(C.POINTER(C.c_uint64))
    # no prior / no proposal / thin < 1 / patient index out of range
    prior = L.Prior()
    bad = [dict(prior=None), dict(prop_sd=0.0), dict(thin=0), dict(pat=np.array([0, 5], dtype=np.int32))]
    for b in bad:
        p_ = b.get("pat", pat)
        rc = lib.gync_mcmc_run_ids(L.dptr(st), L.dptr(lf), 2, None, b.get("prop_sd", 0.1), L.dptr(D), 1, L.dptr(sig),
                                   p_.ctypes.data_as(L.c_int32_p), None if "prior" in b else C.byref(prior), C.byref(op), 10,
                                   b.get("thin", 1), 1, 0, idp, None, None, None, None)
        assert rc == -1, b
    h = C.c_void_p()
    assert lib.gync_em_create(None, 10, 2, C.byref(h)) == -1 and lib.gync_em_create(L.dptr(np.ones((4, 2), order="F")), 0, 2, C.byref(h)) == -1
    assert lib.gync_em_iterate_handle(None, L.dptr(np.ones(4)), 1, None) == -1
    lib.gync_em_destroy(None)                                    # a no-op
    if not torch.cuda.is_available():
        rc = lib.gync_mcmc_run_ids(L.dptr(st), L.dptr(lf), 2, None, 0.1, L.dptr(D), 1, L.dptr(sig), pat.ctypes.data_as(L.c_int32_p),
                                   C.byref(prior), C.byref(op), 10, 1, 1, 0, idp, None, None, None, None)
        assert rc == -2 and b"no CUDA device" in lib.gync_last_error()
        assert lib.gync_em_create(L.dptr(np.ones((4, 2), order="F")), 4, 2, C.byref(h)) == -2


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gync.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                # no import / load / include of anything under oracle/ from the product tree
                assert not re.search(r"import\s+(gync_oracle|c_oracle)|from\s+oracle|libgync_oracle|oracle/_build|"
                                     r"#include\s+\"[^\"]*oracle", src), f


# ---------------------------------------------------------------------------------------- host mirror
def test_api_constants_and_config_defaults(built):
    import gync_b200 as g
    api = g.api
    assert api.sampledinds.size == 82 and np.array_equal(api.sampledinds - 1, o.SAMPLEDINDS)
    assert np.array_equal(api.allparms(api.refparms), api.refallparms)
    c = g.Config()
    assert c.patient.id == "l1" and c.sigma_rho == 0.1 and c.thin == 1 and not c.adapt        # model.jl:51
    assert np.allclose(c.propvar, np.eye(116) * np.log(1.01)) and np.array_equal(c.priorparms, 5 * api.refparms)
    assert np.array_equal(api.init(c), o.init_x())
    assert np.array_equal(np.isnan(g.Lausanne(1).data), np.isnan(o.lausanne(1)))
    assert (~np.isnan(g.Lausanne(1).data)).sum() == 60
    assert len(api.alldatas()) == 53                                                           # notebook: 53 pass minmeas=20
    ca = g.Config(adapt=True)                                                                   # notebooks: production runs adapt
    assert ca.adapt and ca.filename().endswith("atrue.jld")
    v = api.SamplerVariate(ca)
    assert v.m == -1 and v.beta == 0.05 and v.scale == 2.38                                     # model.jl:104-106


def test_api_collapse_duplicates_matches_oracle(built):
    import gync_b200 as g
    rng = np.random.default_rng(3)
    S = rng.random((9, 116))
    S[2] = S[1]
    S[3] = S[1]
    S[5] = S[4]
    c = g.Config()
    s1 = g.Sampling(S[:5], c, None)
    s2 = g.Sampling(S[5:], c, None)          # duplicate run continues across the chain boundary (weightedchain.jl:35-45)
    rows, counts = g.api.collapse_duplicates([s1, s2])
    r0, c0 = o.collapse_duplicates([S[:5], S[5:]])
    assert np.array_equal(rows, r0) and np.array_equal(counts, c0)


def test_shard_bounds():
    from gync_b200.dist import shard_bounds
    for n, w in ((10, 3), (1_000_000, 8), (5, 8)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm of the bench contract) runs without a GPU and prints one JSON line
    with the agreed keys; non-zero ranks print nothing."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, GYNC_BENCH_CPU_BUDGET="1.0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_save_load_roundtrip(built, tmp_path):
    """Checkpoint container (utils.jl:48-49 stand-in): every field of Sampling / Config / variate survives."""
    import gync_b200 as g
    c = g.Config(g.Lausanne(3), sigma_rho=0.2, thin=5, adapt=True,
                 priory0=g.api.GaussianMixture(np.ones((31, 33)), np.ones(33)))
    s = g.new_sampling(c, seed=9)
    s.samples = np.arange(232.0).reshape(2, 116)
    s.variate.logf, s.variate.iteration, s.variate.naccept = -3.5, 10, 4
    s.variate.m, s.variate.Mv = 10, np.arange(116.0)
    s.variate.Mvv, s.variate.SigmaLm = np.eye(116) * 2, np.eye(116) * 3
    path = str(tmp_path / c.filename())
    g.api.save(path, s)
    l = g.api.load(path)
    assert np.array_equal(l.samples, s.samples) and l.config.thin == 5 and l.config.adapt and l.config.sigma_rho == 0.2
    assert l.config.patient.id == "l3" and np.array_equal(np.isnan(l.config.data), np.isnan(c.data))
    v, w = l.variate, s.variate
    assert (v.logf, v.iteration, v.naccept, v.seed, v.m) == (w.logf, w.iteration, w.naccept, w.seed, w.m)
    assert np.array_equal(v.state, w.state) and np.array_equal(v.SigmaL, w.SigmaL) and np.array_equal(v.SigmaLm, w.SigmaLm)
    assert l.config.rtol is None and np.array_equal(l.config.priory0.means, np.ones((31, 33)))


def test_eb_host_mirror_argument_checks(built):
    """Host mirror of src/eb (gync.jl_b200/eb.py): what does not need the device — weights helpers, the pieces that stay
    on the host, and the refusals (no silent CPU fallback for unsupported measurement-error types / autodiff)."""
    import eb_oracle as eo
    from gync_b200 import eb
    assert np.allclose(eb.uniformweights(4), 0.25) and np.allclose(eb.uniformweights([1, 2, 3, 4, 5]), 0.2)   # likelihoodmodel.jl:22-24
    w = np.array([0.2, 0.3, 0.5])
    assert np.array_equal(eb.repeatweights(w, [0] * 6), eo.repeatweights(w, 6))                                # regularizers.jl:16-19
    g = np.array([1.0, -2.0, 0.5])
    assert np.allclose(eb.psi_S(w, g), eo.psi_S(w, g)) and abs(eb.psi_S(w, g).sum() - 1) < 1e-15                # projectsimplex.jl:48-53
    assert np.array_equal(eb.movefromboundary(np.array([0.5, 0.5]), np.array([1.0, 0.0])), [0.75, 0.25])        # gradientascent.jl:19-25
    assert np.allclose(eb.model_measerrors, [12.0, 1.0, 40.0, 1.5])                                            # model.jl:8
    d = eb.MatrixNormalCentered(np.ones((31, 4)))
    assert d == eb.MatrixNormalCentered(np.ones((31, 4))) and hash(d) == hash(eb.MatrixNormalCentered(np.ones((31, 4))))
    with pytest.raises(NotImplementedError):
        eb.likelihoodmat([np.zeros((31, 4))], [np.zeros((31, 4))], "MvNormal")      # the reference's element-wise fallback is not ported
    with pytest.raises(NotImplementedError):
        eb.gradientascent(lambda w_: w_, w, 3, 0.1, autodiff=True)
    with pytest.raises(TypeError):
        eb.projectsimplex_bang([0.5, 0.5])
    m = eb.LikelihoodModel([0], [np.zeros((31, 4))], [], [np.zeros((31, 4))], "not a MatrixNormalCentered")
    with pytest.raises(NotImplementedError):
        m.handle()
