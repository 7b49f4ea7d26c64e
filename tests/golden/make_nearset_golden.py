#!/usr/bin/env python
"""Converged log-likelihoods of two seeded near-set populations (the parity-tail fixtures).

The north-star bound is |llh - llh_converged| <= 1e-8 |llh_converged| for EVERY sample, and the error
distribution of an adaptive solver has a heavy tail, so the -m gpu suite checks
18 432 (sample, patient) cases, not a hand-picked hundred.  Running the C oracle at 1e-12 for them
takes minutes of CPU, so its output is frozen here; the samples themselves are regenerated from the seeds.

    set A: PCG64(99),        2048 samples x Lausanne patients (1, 2, 5)   (the round-1 validation set)
    set B: PCG64(20260101),  4096 samples x Lausanne patients (3, 8, 12)
    set C: PCG64(31337),     4096 samples x Lausanne patients (4, 10, 21) — held out: added after the method and the default
           tolerance had been chosen on A and B, never used to tune anything

x = exp(log(init) + 0.0998 z), z ~ N(0, I_116): one default proposal step from the reference point
(cfg 2 "near").
    set W: PCG64(777), 512 samples of the WIDE set (eb/gync.jl:59: parameters uniform on (0, 5 ref), y0 ~ GMM prior,
           T = 30) x Lausanne patient 1; 77 % of them do not solve (-Inf) — the failure SET is part of the fixture.
Run in the build container (~10 min on 8 cores):
    python tests/golden/make_nearset_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

SETS = {"A": (99, 2048, (0, 1, 4)), "B": (20260101, 4096, (2, 7, 11)),
        # HELD OUT: drawn and solved after the solver, its tolerance and every other knob had been fixed on A and B
        "C": (31337, 4096, (3, 9, 20))}


def wide_set(seed, n):
    import gync_oracle as o
    g = np.load(os.path.join(HERE, "solution_golden.npz"))
    means, stds = g["gmm_means"], g["gmm_stds"]
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.array([np.concatenate([o.REFPARMS * rng.random(82) * 5, means[rng.integers(31)] + stds * rng.standard_normal(33),
                                     [30.0]]) for _ in range(n)])


def near_set(seed, n):
    import gync_oracle as o
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.exp(np.log(o.init_x())[None, :] + 0.0998 * rng.standard_normal((n, 116)))


if __name__ == "__main__":
    import c_oracle as co
    import gync_oracle as o
    la = o.patients()[0]
    out = {}
    for name, (seed, n, pats) in SETS.items():
        X = near_set(seed, n)
        LL = co.llh_batch(X, [la[i] for i in pats], [0.1] * len(pats), 1, 1e-12, 1e-12)
        assert np.isfinite(LL).all()
        out["LL_" + name] = LL
        out["seed_" + name] = seed
        out["pats_" + name] = np.array(pats)
        print(name, LL.shape, float(LL.mean()))
    Xw = wide_set(777, 512)
    out["LL_W"] = co.llh_batch(Xw, [la[0]], [0.1], 1, 1e-12, 1e-12)[:, 0]
    out["seed_W"] = 777
    print("W finite fraction", float(np.isfinite(out["LL_W"]).mean()))
    np.savez_compressed(os.path.join(HERE, "nearset_llh_golden.npz"), sigma=0.1, oracle_rtol=1e-12, **out)
