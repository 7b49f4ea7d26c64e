"""CPU ORACLE (test infrastructure, NOT product code) for the GynC.jl hot path.

A plain numpy/scipy restatement of the reference's arithmetic, every function citing
the reference file:line it follows (paths relative to the reference repo root).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

PARITY UNPINNED at the solver level: the reference integrates with SUNDIALS CVODE
(Sundials.jl, REQUIRE:9, no version pin, not vendored) and its tests hold no numeric
golden vectors on this path (test/gync.jl:14 pins a shape only).  What IS pinned:
  * the RHS  — tests/golden/rhs_golden.npz was produced by mechanically transliterating
    the reference's own source text of gyncycle_rhs! at generation time
    (tests/golden/make_rhs_golden.py) and this module is checked against it;
  * refy0 / refallparms / patient tables — data extracted by tools/extract_reference_data.py.
Trajectories / log-likelihoods are pinned against the *converged* solution of the same
ODE (two independent integrators agreeing to <=1e-10, tests/golden/make_solution_golden.py),
which is the reference's own path run with tight `reltol/abstol` kwargs (gyncycle.jl:8).
"""
from __future__ import annotations

import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_HERE, "..", "gync.jl_b200", "data")

with open(os.path.join(_DATA, "reference_constants.json")) as _f:
    _C = json.load(_f)

#: src/gync/data/refy0.jl, src/gync/data/refparms.jl (gyncycle.jl:1-2)
REFY0 = np.array(_C["refy0"], dtype=np.float64)
REFALLPARMS = np.array(_C["refallparms"], dtype=np.float64)

#: model.jl:1  (1-based in the reference; 0-based here)
MEASUREDINDS = np.array([2, 7, 24, 25]) - 1
#: model.jl:3
HILLINDS = np.array([4, 6, 10, 18, 20, 22, 26, 33, 36, 39, 43, 47, 49, 52, 55, 59, 65, 95, 98, 101, 103]) - 1
#: model.jl:4  deleteat!(collect(1:103), hillinds)
SAMPLEDINDS = np.array([i for i in range(103) if i not in set(HILLINDS.tolist())])
assert SAMPLEDINDS.size == 82
#: model.jl:6
REFPARMS = REFALLPARMS[SAMPLEDINDS]
#: model.jl:10
MEASMAGNITUDE = np.array([120.0, 10.0, 400.0, 15.0])
#: model.jl:14-16   eye(116) * log(1 + 0.1^2)
DEFAULT_PROPVAR_DIAG = np.full(116, np.log(1.0 + 0.1 ** 2))

NX, NY, NP, NS = 116, 33, 114, 82


def patients():
    """Patient tables as produced by data.jl:10-36 (Lausanne) and :41-58 (Pfizer)."""
    z = np.load(os.path.join(_DATA, "patients.npz"))
    return z["lausanne"], z["pfizer"]


def lausanne(i: int) -> np.ndarray:
    """Lausanne(id).data — 31x4, NaN = missing (data.jl:8, 1-based id)."""
    return patients()[0][i - 1].copy()


# --------------------------------------------------------------------------- RHS
def _hplus(s, t, n):
    """gyncycle.jl:41-44"""
    q = (s / t) ** n
    return q / (1.0 + q)


def _hminus(s, t, n):
    """gyncycle.jl:46"""
    return 1.0 / (1.0 + (s / t) ** n)


def gyncycle_rhs(y, p):
    """dy/dt of the 33 species — restates gyncycle_rhs! (gyncycle.jl:50-244).

    y: (...,33), p: (...,114) broadcastable; returns (...,33).
    Species order gyncycle.jl:60-92.  Hill exponents are literals (gyncycle.jl:99-182).
    Negative base of the real power -> NaN (Julia raises DomainError -> NaN trajectory,
    gyncycle.jl:18-19).
    """
    y = np.asarray(y, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    Y = lambda i: y[..., i - 1]     # 1-based accessors keep the equation numbers readable
    P = lambda i: p[..., i - 1]
    (LH_pit, LH_bl, R_LH, LH_R, R_LH_des, FSH_pit, FSH_bl, R_FSH, FSH_R, R_FSH_des, s_, AF1, AF2, AF3, AF4,
     PrF, OvF, Sc1, Sc2, Lut1, Lut2, Lut3, Lut4, E2, P4, IhA, IhB, IhA_e, G, R_G_a, R_G_i, G_R_a, G_R_i) = range(1, 34)
    f = np.empty(np.broadcast(y[..., 0], p[..., 0]).shape + (33,), dtype=np.float64)

    with np.errstate(all="ignore"):
        # eq 29a/b  (:99-100)
        y_freq = P(93) * _hminus(Y(P4), P(94), 2) * (1.0 + P(96) * _hplus(Y(E2), P(97), 10))
        y_mass = P(99) * (_hplus(Y(E2), P(100), 2) + _hminus(Y(E2), P(102), 1))
        # eq 1-2  (:103-108)
        syn_lh = (P(1) + P(2) * _hplus(Y(E2), P(3), 10)) * _hminus(Y(P4), P(5), 1)
        rel_lh = P(7) + P(8) * _hplus(Y(G_R_a), P(9), 5)
        f[..., LH_pit - 1] = syn_lh - rel_lh * Y(LH_pit)
        f[..., LH_bl - 1] = rel_lh * Y(LH_pit) / P(11) - (P(12) * Y(R_LH) + P(13)) * Y(LH_bl)
        # eq 3-5  (:111-117)
        f[..., R_LH - 1] = P(14) * Y(R_LH_des) - P(12) * Y(LH_bl) * Y(R_LH)
        f[..., LH_R - 1] = P(12) * Y(LH_bl) * Y(R_LH) - P(15) * Y(LH_R)
        f[..., R_LH_des - 1] = P(15) * Y(LH_R) - P(14) * Y(R_LH_des)
        # eq 6-7  (:120-125)
        syn_fsh = P(16) / (1.0 + (Y(IhA_e) / P(17)) ** 5 + (Y(IhB) / P(19)) ** 3) * _hminus(y_freq, P(21), 3)
        rel_fsh = P(23) + P(24) * _hplus(Y(G_R_a), P(25), 3)
        f[..., FSH_pit - 1] = syn_fsh - rel_fsh * Y(FSH_pit)
        f[..., FSH_bl - 1] = rel_fsh * Y(FSH_pit) / P(11) - (P(27) * Y(R_FSH) + P(28)) * Y(FSH_bl)
        # eq 8-10  (:128-134)
        f[..., R_FSH - 1] = P(29) * Y(R_FSH_des) - P(27) * Y(FSH_bl) * Y(R_FSH)
        f[..., FSH_R - 1] = P(27) * Y(FSH_bl) * Y(R_FSH) - P(30) * Y(FSH_R)
        f[..., R_FSH_des - 1] = P(30) * Y(FSH_R) - P(29) * Y(R_FSH_des)
        # eq 11-12  (:137-140)
        f[..., s_ - 1] = P(31) * _hplus(Y(FSH_bl), P(32), 5) - P(34) * _hplus(Y(P4), P(35), 5) * Y(s_)
        f[..., AF1 - 1] = P(37) * _hplus(Y(FSH_R), P(38), 3) - P(40) * Y(FSH_R) * Y(AF1)
        # the hoisted real power (:144)
        r = Y(LH_R) / P(42)
        rpow = np.where(r >= 0, np.abs(r) ** 3.689, np.nan)
        # eq 13-17  (:147-167)
        f[..., AF2 - 1] = P(40) * Y(FSH_R) * Y(AF1) - P(41) * rpow * Y(s_) * Y(AF2)
        f[..., AF3 - 1] = (P(41) * rpow * Y(s_) * Y(AF2)
                           + P(44) * Y(FSH_R) * Y(AF3) * (1 - Y(AF3) / P(45))
                           - P(46) * r ** 5 * Y(s_) * Y(AF3))
        f[..., AF4 - 1] = (P(46) * r ** 5 * Y(s_) * Y(AF3)
                           + P(48) * r ** 2 * Y(AF4) * (1.0 - Y(AF4) / P(45))
                           - P(50) * r * Y(s_) * Y(AF4))
        f[..., PrF - 1] = P(50) * r * Y(s_) * Y(AF4) - P(51) * r ** 6 * Y(s_) * Y(PrF)
        f[..., OvF - 1] = P(53) * r ** 6 * Y(s_) * _hplus(Y(PrF), P(54), 10) - P(56) * Y(OvF)
        # eq 18-23  (:170-185)
        f[..., Sc1 - 1] = P(57) * _hplus(Y(OvF), P(58), 10) - P(60) * Y(Sc1)
        f[..., Sc2 - 1] = P(60) * Y(Sc1) - P(61) * Y(Sc2)
        lut = 1.0 + P(63) * _hplus(Y(G_R_a), P(64), 5)
        f[..., Lut1 - 1] = P(61) * Y(Sc2) - P(62) * lut * Y(Lut1)
        f[..., Lut2 - 1] = P(62) * Y(Lut1) - P(66) * lut * Y(Lut2)
        f[..., Lut3 - 1] = P(66) * Y(Lut2) - P(67) * lut * Y(Lut3)
        f[..., Lut4 - 1] = P(67) * Y(Lut3) - P(68) * lut * Y(Lut4)
        # eq 24-28  (:188-215)
        f[..., E2 - 1] = (P(69) + P(70) * Y(AF2) + P(71) * Y(LH_bl) * Y(AF3) + P(72) * Y(AF4)
                          + P(73) * Y(LH_bl) * Y(PrF) + P(74) * Y(Lut1) + P(75) * Y(Lut4) - P(76) * Y(E2))
        f[..., P4 - 1] = P(77) + P(78) * Y(Lut4) - P(79) * Y(P4)
        f[..., IhA - 1] = (P(80) + P(81) * Y(PrF) + P(82) * Y(Sc1) + P(83) * Y(Lut1) + P(84) * Y(Lut2)
                           + P(85) * Y(Lut3) + P(86) * Y(Lut4) - P(87) * Y(IhA))
        f[..., IhB - 1] = P(88) + P(89) * Y(AF2) + P(90) * Y(Sc2) - P(91) * Y(IhB)
        f[..., IhA_e - 1] = P(87) * Y(IhA) - P(92) * Y(IhA_e)
        # eq 29-33  (:218-240)
        bind = P(104) * Y(G) * Y(R_G_a)
        f[..., G - 1] = y_mass * y_freq - bind + P(105) * Y(G_R_a) - P(106) * Y(G)
        f[..., R_G_a - 1] = P(105) * Y(G_R_a) - bind - P(107) * Y(R_G_a) + P(108) * Y(R_G_i)
        f[..., R_G_i - 1] = (P(109) + P(107) * Y(R_G_a) - P(108) * Y(R_G_i) + P(110) * Y(G_R_i) - P(111) * Y(R_G_i))
        f[..., G_R_a - 1] = bind - P(105) * Y(G_R_a) - P(112) * Y(G_R_a) + P(113) * Y(G_R_i)
        f[..., G_R_i - 1] = P(112) * Y(G_R_a) - P(113) * Y(G_R_i) - P(110) * Y(G_R_i) - P(114) * Y(G_R_i)
    return f


def numjac(y, p, rel=1e-7):
    """Central-difference Jacobian of gyncycle_rhs (used only to check analytic Jacobians)."""
    y = np.asarray(y, dtype=np.float64)
    J = np.empty((33, 33))
    for j in range(33):
        h = rel * max(abs(y[j]), 1e-8)
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        J[:, j] = (gyncycle_rhs(yp, p) - gyncycle_rhs(ym, p)) / (yp[j] - ym[j])
    return J


# ---------------------------------------------------------------- model.jl glue
def allparms(parms):
    """model.jl:20-24 — scatter the 82 sampled parameters into a copy of refallparms."""
    parms = np.asarray(parms, dtype=np.float64)
    p = np.broadcast_to(REFALLPARMS, parms.shape[:-1] + (114,)).copy()
    p[..., SAMPLEDINDS] = parms
    return p


def split(x):
    """model.jl:82-84 — parms(x)=x[1:82], y0(x)=x[83:115], period(x)=x[116]."""
    x = np.asarray(x, dtype=np.float64)
    return x[..., :82], x[..., 82:115], x[..., 115]


def init_x():
    """model.jl:94 with Config() defaults (model.jl:51): [refparms; refy0; 28]."""
    return np.concatenate([REFPARMS, REFY0, [28.0]])


def meastimes(days, periods, T):
    """model.jl:126 — vcat([(1:days) + T*p for p in 0:periods-1]...)."""
    return np.concatenate([np.arange(1, days + 1) + T * q for q in range(periods)]).astype(np.float64)


def _fast_rhs(p):
    return lambda _t, y: gyncycle_rhs(y, p)


def solve_at(y0, p, t, rtol, atol, method="Radau"):
    """Converged-mode trajectory: integrate piecewise, landing EXACTLY on each output time
    (no dense-output interpolation error).  Returns (nT,33) or all-NaN on failure."""
    from scipy.integrate import solve_ivp
    y0 = np.asarray(y0, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    out = np.full((t.size, 33), np.nan)
    if not (np.all(np.isfinite(y0)) and np.all(np.isfinite(p))):
        return out
    fun = _fast_rhs(p)
    jac = (lambda _t, y: analytic_jac(y, p))
    y = y0.copy()
    out[0] = y
    first_step = None
    with np.errstate(all="ignore"):
        for k in range(1, t.size):
            if t[k] > t[k - 1]:
                kw = {"jac": jac} if method in ("Radau", "BDF") else {}
                try:
                    s = solve_ivp(fun, (t[k - 1], t[k]), y, method=method, rtol=rtol, atol=atol,
                                  first_step=first_step, **kw)
                except Exception:
                    return np.full((t.size, 33), np.nan)
                if not s.success or not np.all(np.isfinite(s.y[:, -1])):
                    return np.full((t.size, 33), np.nan)
                y = s.y[:, -1]
                if s.t.size > 2:
                    first_step = min(s.t[-2] - s.t[-3], 0.5 * (t[min(k + 1, t.size - 1)] - t[k]) or 1.0)
                    first_step = first_step if first_step > 0 else None
            out[k] = y
    return out


def analytic_jac(y, p):
    """Dense 33x33 Jacobian by complex-step differentiation of the restated RHS
    (exact to rounding for this analytic RHS; independent of the product's symbolic Jacobian)."""
    y = np.asarray(y, dtype=np.float64)
    J = np.empty((33, 33))
    h = 1e-30
    yc = y.astype(np.complex128)
    for j in range(33):
        yy = yc.copy()
        yy[j] += 1j * h
        J[:, j] = _rhs_complex(yy, p).imag / h
    return J


def _rhs_complex(y, p):
    """gyncycle_rhs evaluated on complex y (same arithmetic; real power via exp/log)."""
    Y = lambda i: y[i - 1]
    P = lambda i: p[i - 1]
    hp = lambda s, t, n: ((s / t) ** n) / (1.0 + (s / t) ** n)
    hm = lambda s, t, n: 1.0 / (1.0 + (s / t) ** n)
    f = np.empty(33, dtype=np.complex128)
    y_freq = P(93) * hm(Y(25), P(94), 2) * (1.0 + P(96) * hp(Y(24), P(97), 10))
    y_mass = P(99) * (hp(Y(24), P(100), 2) + hm(Y(24), P(102), 1))
    syn_lh = (P(1) + P(2) * hp(Y(24), P(3), 10)) * hm(Y(25), P(5), 1)
    rel_lh = P(7) + P(8) * hp(Y(32), P(9), 5)
    f[0] = syn_lh - rel_lh * Y(1)
    f[1] = rel_lh * Y(1) / P(11) - (P(12) * Y(3) + P(13)) * Y(2)
    f[2] = P(14) * Y(5) - P(12) * Y(2) * Y(3)
    f[3] = P(12) * Y(2) * Y(3) - P(15) * Y(4)
    f[4] = P(15) * Y(4) - P(14) * Y(5)
    syn_fsh = P(16) / (1.0 + (Y(28) / P(17)) ** 5 + (Y(27) / P(19)) ** 3) * hm(y_freq, P(21), 3)
    rel_fsh = P(23) + P(24) * hp(Y(32), P(25), 3)
    f[5] = syn_fsh - rel_fsh * Y(6)
    f[6] = rel_fsh * Y(6) / P(11) - (P(27) * Y(8) + P(28)) * Y(7)
    f[7] = P(29) * Y(10) - P(27) * Y(7) * Y(8)
    f[8] = P(27) * Y(7) * Y(8) - P(30) * Y(9)
    f[9] = P(30) * Y(9) - P(29) * Y(10)
    f[10] = P(31) * hp(Y(7), P(32), 5) - P(34) * hp(Y(25), P(35), 5) * Y(11)
    f[11] = P(37) * hp(Y(9), P(38), 3) - P(40) * Y(9) * Y(12)
    r = Y(4) / P(42)
    rpow = np.exp(3.689 * np.log(r))
    f[12] = P(40) * Y(9) * Y(12) - P(41) * rpow * Y(11) * Y(13)
    f[13] = (P(41) * rpow * Y(11) * Y(13) + P(44) * Y(9) * Y(14) * (1 - Y(14) / P(45))
             - P(46) * r ** 5 * Y(11) * Y(14))
    f[14] = (P(46) * r ** 5 * Y(11) * Y(14) + P(48) * r ** 2 * Y(15) * (1.0 - Y(15) / P(45))
             - P(50) * r * Y(11) * Y(15))
    f[15] = P(50) * r * Y(11) * Y(15) - P(51) * r ** 6 * Y(11) * Y(16)
    f[16] = P(53) * r ** 6 * Y(11) * hp(Y(16), P(54), 10) - P(56) * Y(17)
    f[17] = P(57) * hp(Y(17), P(58), 10) - P(60) * Y(18)
    f[18] = P(60) * Y(18) - P(61) * Y(19)
    lut = 1.0 + P(63) * hp(Y(32), P(64), 5)
    f[19] = P(61) * Y(19) - P(62) * lut * Y(20)
    f[20] = P(62) * Y(20) - P(66) * lut * Y(21)
    f[21] = P(66) * Y(21) - P(67) * lut * Y(22)
    f[22] = P(67) * Y(22) - P(68) * lut * Y(23)
    f[23] = (P(69) + P(70) * Y(13) + P(71) * Y(2) * Y(14) + P(72) * Y(15) + P(73) * Y(2) * Y(16)
             + P(74) * Y(20) + P(75) * Y(23) - P(76) * Y(24))
    f[24] = P(77) + P(78) * Y(23) - P(79) * Y(25)
    f[25] = (P(80) + P(81) * Y(16) + P(82) * Y(18) + P(83) * Y(20) + P(84) * Y(21) + P(85) * Y(22)
             + P(86) * Y(23) - P(87) * Y(26))
    f[26] = P(88) + P(89) * Y(13) + P(90) * Y(19) - P(91) * Y(27)
    f[27] = P(87) * Y(26) - P(92) * Y(28)
    bind = P(104) * Y(29) * Y(30)
    f[28] = y_mass * y_freq - bind + P(105) * Y(32) - P(106) * Y(29)
    f[29] = P(105) * Y(32) - bind - P(107) * Y(30) + P(108) * Y(31)
    f[30] = P(109) + P(107) * Y(30) - P(108) * Y(31) + P(110) * Y(33) - P(111) * Y(31)
    f[31] = bind - P(105) * Y(32) - P(112) * Y(32) + P(113) * Y(33)
    f[32] = P(112) * Y(32) - P(113) * Y(33) - P(110) * Y(33) - P(114) * Y(33)
    return f


def forwardsol(x, tspan, rtol=1e-4, atol=1e-4, method="BDF", exact_times=False):
    """forwardsol(x,tspan) = gync(y0(x), allparms(parms(x)), tspan)  (model.jl:90).

    exact_times=False: one integration with dense-output interpolation (CVODE CV_NORMAL
    semantics of the reference path); exact_times=True: converged mode, step onto each t."""
    parms, y0, _ = split(x)
    p = allparms(parms)
    t = np.asarray(tspan, dtype=np.float64)
    if exact_times:
        return solve_at(y0, p, t, rtol, atol, method)
    from scipy.integrate import solve_ivp
    nan = np.full((t.size, 33), np.nan)
    if not (np.all(np.isfinite(y0)) and np.all(np.isfinite(p))):
        return nan
    try:
        with np.errstate(all="ignore"):
            kw = {"jac": (lambda _t, yy: analytic_jac(yy, p))} if method in ("Radau", "BDF") else {}
            s = solve_ivp(_fast_rhs(p), (t[0], t[-1]), y0, method=method, t_eval=t, rtol=rtol, atol=atol, **kw)
    except Exception:
        return nan
    if not s.success or s.y.shape[1] != t.size:
        return nan
    return s.y.T.copy()


# ------------------------------------------------------------------ likelihood
def llh_measerror(error, sigma_rho):
    """model.jl:158-163 — -sum((err/(mag*sigma))^2)/2 over non-NaN; all-NaN -> -Inf."""
    scaled = np.asarray(error) / (MEASMAGNITUDE[None, :] * sigma_rho)
    nonnan = ~np.isnan(scaled)
    if not nonnan.any():
        return -np.inf
    return -np.sum(scaled[nonnan] ** 2) / 2.0


def llh_from_traj(y_meas, data, sigma_rho, periods=2):
    """Tail of llh (model.jl:138-155) given the 62x4 measured-species trajectory in
    [period0; period1] order."""
    if np.isnan(y_meas).any():
        return -np.inf
    nd = data.shape[0]
    l = 0.0
    for q in range(periods):
        l += llh_measerror(data - y_meas[q * nd:(q + 1) * nd, :], sigma_rho)
    return l


def traj_at_meastimes(x, rtol, atol, method="Radau", ndata=31, periods=2, exact_times=True):
    """Head of llh (model.jl:129-136): times, sortperm, solve, invperm, measured species.
    Returns (62,4)."""
    T = float(np.asarray(x)[115])
    t = meastimes(ndata, periods, T)
    perm = np.argsort(t, kind="stable")
    sol = forwardsol(x, t[perm], rtol, atol, method, exact_times=exact_times)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return sol[inv][:, MEASUREDINDS]


def llh(x, data, sigma_rho=0.1, rtol=1e-4, atol=1e-4, method="BDF", exact_times=False):
    """llh(c,x,periods=2)  (model.jl:128-155)."""
    ym = traj_at_meastimes(x, rtol, atol, method, data.shape[0], 2, exact_times)
    return llh_from_traj(ym, data, sigma_rho)


# ----------------------------------------------------------------------- prior
def referencesolution(sol=None):
    """gyncycle.jl:30-38 — gync(refy0, allparms(refparms), 0:30) with non-positive entries of
    each column replaced by that column's minimum positive entry.  The reference integrates
    at reltol=abstol=1e-4; pass `sol` to apply the fix-up to a given 31x33 trajectory."""
    if sol is None:
        sol = forwardsol(init_x(), np.arange(0.0, 31.0), 1e-4, 1e-4, "BDF")
    sol = np.array(sol, dtype=np.float64)
    for j in range(sol.shape[1]):
        small = sol[:, j] <= 0
        if small.any():
            sol[small, j] = sol[~small, j].min()
    return sol


def gmm_from_refsol(refsol, stdfactor=1.0):
    """gaussianmixture(y, stdfactor) (model.jl:73-77): uniform mixture of 31 diagonal normals
    centred on the rows of y; MvNormal(yt, stds) takes stds = column std (corrected, n-1)."""
    stds = refsol.std(axis=0, ddof=1) * stdfactor
    return refsol.copy(), stds


def gmm_logpdf(y0, means, stds):
    """Distributions.logpdf(MixtureModel(MvNormal(mean_i, stds)), y0) with uniform weights."""
    z = (y0[None, :] - means) / stds[None, :]
    lp = -0.5 * np.sum(z * z, axis=1) - np.sum(np.log(stds)) - 0.5 * means.shape[1] * np.log(2 * np.pi)
    m = lp.max()
    if not np.isfinite(m):
        return -np.inf
    return m + np.log(np.sum(np.exp(lp - m))) - np.log(means.shape[0])


def logprior(x, means, stds, upper=None, mu_T=28.9, sd_T=3.4):
    """logprior(c,x) (model.jl:111-117): GMM(y0) + sum_82 Truncated(Flat,0,5*refparms_i) + Normal(28.9,3.4)(T).
    Truncated(Flat(),0,a): constant (taken as 0 here — it cancels everywhere on the hot path,
    SURVEY Appendix D) inside [0,a], -Inf outside."""
    parms, y0, T = split(x)
    upper = 5.0 * REFPARMS if upper is None else upper
    if np.any(parms < 0) or np.any(parms > upper):
        return -np.inf
    l = gmm_logpdf(y0, means, stds)
    l += -0.5 * ((T - mu_T) / sd_T) ** 2 - np.log(sd_T) - 0.5 * np.log(2 * np.pi)
    return float(l)


def logpost(x, data, sigma_rho, means, stds, **kw):
    """model.jl:119-124."""
    l = logprior(x, means, stds)
    if l == -np.inf:
        return l
    return l + llh(x, data, sigma_rho, **kw)


# --------------------------------------------------------------- Metropolis step
def metropolis_step(state, cur_logf, chol, z, u, logf):
    """One non-adaptive Mamba AMM step (call site sampling.jl:72, model.jl:102-106):
    x' = x + SigmaL*z ; accept iff u < exp(logf(x') - logf(x)).  `logf(x)=logpost(exp(x))+sum(x)`.
    Randomness (z ~ N(0,I_116), u ~ U(0,1)) is injected so the step is deterministic."""
    prop = state + chol @ z
    lp = logf(prop)
    if np.log(u) < lp - cur_logf:
        return prop, lp, True
    return state, cur_logf, False


def amm_new_tune(state, SigmaL, beta=0.05, scale=2.38):
    """AMMTune after setadapt!(v, true) on a fresh variate: m = 0, Mv = v, Mvv = v v',
    SigmaLm = SigmaL until the running covariance has full rank (a zero factor would freeze the chain at m > 2d)."""
    v = np.asarray(state, dtype=np.float64)
    return {"m": 0, "Mv": v.copy(), "Mvv": np.outer(v, v), "SigmaL": np.asarray(SigmaL, dtype=np.float64),
            "SigmaLm": np.array(SigmaL, dtype=np.float64), "beta": beta, "scale": scale}


def chol_fullrank(S):
    """Lower Cholesky factor of S, or None when S is not numerically of full rank — the LAPACK dpstrf stopping rule
    (pivot <= n eps max(diag)) that `rank(cholfact(S, :U, Val{true})) == n` tests in Mamba, on the unpivoted
    factorisation (same Sigma = L L'; Mamba uses the pivoted square root P L — same proposal distribution)."""
    S = np.array(S, dtype=np.float64)
    n = S.shape[0]
    tol = n * np.finfo(float).eps * np.max(np.diag(S))
    L = np.zeros_like(S)
    for k in range(n):
        piv = S[k, k] - np.dot(L[k, :k], L[k, :k])
        if not piv > tol:
            return None
        L[k, k] = np.sqrt(piv)
        L[k + 1:, k] = (S[k + 1:, k] - L[k + 1:, :k] @ L[k, :k]) / L[k, k]
    return L


def amm_step(state, cur_logf, tune, z, u, umix, logf):
    """One ADAPTIVE Mamba AMM step (Mamba.sample!(v, adapt=true); SURVEY Appendix D — recalled third-party behaviour,
    parity unpinned; call sites model.jl:104-106, sampling.jl:72).  Mutates `tune`."""
    d = state.size
    tune["m"] += 1
    m = tune["m"]
    F = tune["SigmaLm"] if (m > 2 * d and umix >= tune["beta"]) else tune["SigmaL"]
    prop = state + F @ z
    lp = logf(prop)
    acc = bool(np.log(u) < lp - cur_logf)
    if acc:
        state, cur_logf = prop, lp
    p = m / (m + 1.0)
    tune["Mv"] = p * tune["Mv"] + (1 - p) * state
    tune["Mvv"] = p * tune["Mvv"] + (1 - p) * np.outer(state, state)
    Sigma = (tune["scale"] ** 2 / d / p) * (tune["Mvv"] - np.outer(tune["Mv"], tune["Mv"]))
    L = chol_fullrank(Sigma) if m >= d else None
    if L is not None:
        tune["SigmaLm"] = L
    return state, cur_logf, acc


# --------------------------------------------------------------- WeightedChain/EM
def weightedchain_normalize(logprior_k, LL, counts):
    """weightedchain.jl:59-67 — exp-normalise, column-normalise, weights and density."""
    prior = np.exp(logprior_k - logprior_k.max())
    lhs = np.exp(LL - LL.max(axis=0, keepdims=True))
    lhs = lhs / lhs.sum(axis=0, keepdims=True)
    weights = counts / counts.sum()
    density = lhs.mean(axis=1) * prior
    density = density / density.sum()
    return lhs, weights, density


def collapse_duplicates(samples_list, burnin=0):
    """weightedchain.jl:31-46 — collapse CONSECUTIVE duplicate rows across the concatenated
    samplings; `curr` starts as a zero row."""
    curr = np.zeros(samples_list[0].shape[1])
    rows, counts = [], []
    for S in samples_list:
        for i in range(burnin, S.shape[0]):
            if np.array_equal(S[i], curr) and counts:
                counts[-1] += 1
            elif np.array_equal(S[i], curr):
                # reference would index counts[end] of an empty vector -> error; unreachable for
                # positive samples (curr starts as zeros)
                raise IndexError("first row equals the zero vector")
            else:
                curr = S[i]
                rows.append(curr.copy())
                counts.append(1)
    return np.array(rows), np.array(counts, dtype=np.int64)


def emiteration(w, L):
    """emiteration!(w,L) (em.jl:27-41), returning the new w:
    norms = L'w ; w[k] <- w[k]/M * sum_m L[k,m]/norms[m].  Sequential two-pass order."""
    w = np.array(w, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    K, M = L.shape
    norms = L.T @ w
    with np.errstate(all="ignore"):
        s = np.zeros(K)
        for m in range(M):            # same m-order accumulation as the reference loop
            s += L[:, m] / norms[m]
        return w / M * s


def emiterations(w0, L, niter):
    """em.jl:5 — collect(take(iterate(w->emiteration(w,L), w0), niter)): w0 first, niter entries."""
    out = [np.array(w0, dtype=np.float64)]
    for _ in range(niter - 1):
        out.append(emiteration(out[-1], L))
    return out
