#!/usr/bin/env python
"""Freeze converged trajectories / log-likelihoods / prior / EM vectors produced by the ORACLE.

Nothing in the reference pins these numerically (shapes only), so they are generated
here by two independent integrators — the C oracle's extrapolation code (oracle/gync_oracle.c,
converged mode) and scipy's Radau IIA on the same restated RHS — which must agree to <=1e-9
relative in log-likelihood before anything is written.  Run in the build container:
    python tests/golden/make_solution_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import c_oracle as co  # noqa: E402
import gync_oracle as o  # noqa: E402
from scipy.integrate import solve_ivp  # noqa: E402


def radau_traj(x, rtol=1e-12, atol=1e-14):
    parms, y0, T = o.split(x)
    p = o.allparms(parms)
    t = o.meastimes(31, 2, T)
    perm = np.argsort(t, kind="stable")
    ts = t[perm]
    L = co.lib()
    J = np.empty((33, 33))

    def jac(_t, y):
        L.orc_jac_exact(co.P(np.ascontiguousarray(y)), co.P(p), co.P(J))
        return J.copy()
    out = np.empty((62, 33))
    y = y0.copy()
    out[0] = y
    for k in range(1, 62):
        if ts[k] > ts[k - 1]:
            s = solve_ivp(lambda _t, yy: co.rhs(yy, p), (ts[k - 1], ts[k]), y, method="Radau", rtol=rtol, atol=atol, jac=jac)
            assert s.success
            y = s.y[:, -1]
        out[k] = y
    inv = np.empty_like(perm)
    inv[perm] = np.arange(62)
    return out[inv][:, o.MEASUREDINDS]


if __name__ == "__main__":
    rng = np.random.Generator(np.random.PCG64(20160911 + 2))
    x0 = o.init_x()
    near = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((32, 116)))   # cfg 2 "near" set
    near[0] = x0
    near[1, 115] = 28.0          # integer period: duplicated measurement times (model.jl:94,126)
    near[2, 115] = 30.5
    # converged reference solution -> y0 prior (model.jl:69, gyncycle.jl:30-38)
    refsol = co.gync(o.REFY0, o.allparms(o.REFPARMS), np.arange(0.0, 31.0), mode=1, rtol=1e-12, atol=1e-12)
    refsol = o.referencesolution(refsol)
    means, stds = o.gmm_from_refsol(refsol)
    # "wide" set (eb/gync.jl:59): refparms .* 5u, y0 ~ GMM prior, T = 30
    wide = []
    while len(wide) < 6:
        comp = rng.integers(31)
        x = np.concatenate([o.REFPARMS * rng.random(82) * 5, means[comp] + stds * rng.standard_normal(33), [30.0]])
        wide.append(x)
    X = np.vstack([near, np.array(wide)])
    # patients: Lausanne(1) + synthetic (simulate a near sample, add noise, Lausanne missingness; eb/gync.jl:20-22)
    la = o.patients()[0]
    datas = [la[0]]
    sig = 0.1
    for m in range(1, 4):
        tr = co.llh_batch(near[3 + m][None, :], [la[0]], [sig], 1, 1e-10, 1e-10, want_traj=True)[1][0]
        d = tr[:31] + rng.standard_normal((31, 4)) * (sig * o.MEASMAGNITUDE)[None, :]
        d[np.isnan(la[m])] = np.nan
        datas.append(d)
    t0 = time.time()
    LL, traj = co.llh_batch(X, datas, [sig] * 4, 1, 1e-12, 1e-12, want_traj=True)
    print("C oracle converged: %.1fs" % (time.time() - t0), "non-finite rows:", np.flatnonzero(~np.isfinite(LL[:, 0])))
    # independent cross-check with Radau on 5 samples
    chk = [0, 1, 2, 7, 19]
    for i in chk:
        t0 = time.time()
        tr = radau_traj(X[i])
        for m in range(4):
            l = o.llh_from_traj(tr, datas[m], sig)
            rel = abs(l - LL[i, m]) / abs(l)
            assert rel < 1e-9, (i, m, l, LL[i, m], rel)
        rt = np.max(np.abs(tr - traj[i]) / np.abs(tr))
        print("sample %d: Radau vs extrapolation  llh rel %.1e  traj rel %.1e  (%.0fs)" % (i, rel, rt, time.time() - t0))
        assert rt < 1e-8
    # reference-tolerance mode beside it (BDF 1e-4, the CVODE stand-in)
    LL4 = co.llh_batch(X, datas[:1], [sig], 0, 1e-4, 1e-4)
    # prior
    LP = np.array([o.logprior(x, means, stds) for x in X])
    # EM
    K, M = 2000, 40
    Lm = np.exp(-0.5 * rng.standard_normal((K, M)) ** 2 * 8)
    Lm /= Lm.sum(0, keepdims=True)
    w = np.full(K, 1.0 / K)
    ws = {}
    for it in range(1, 101):
        w = o.emiteration(w, Lm)
        if it in (1, 10, 100):
            ws[it] = w.copy()
    np.savez_compressed(os.path.join(HERE, "solution_golden.npz"), X=X, data=np.stack(datas, axis=2), sigma=sig, LL=LL,
                        traj=traj, LL_bdf1e4=LL4, refsol=refsol, gmm_means=means, gmm_stds=stds, logprior=LP,
                        em_L=Lm, em_w1=ws[1], em_w10=ws[10], em_w100=ws[100], radau_checked=np.array(chk))
    print("wrote solution_golden.npz; llh(ref point, Lausanne 1) = %.12f; BDF1e-4: %.6f" % (LL[0, 0], LL4[0, 0]))
