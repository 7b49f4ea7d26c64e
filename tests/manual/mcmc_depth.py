"""Speculation depth sweep (GYNC_MCMC_SPEC) for many-chain Metropolis: chain-iterations/s as a function of K."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o
L.check(L.load().gync_init(0))
la = o.patients()[0]
C_, iters = int(os.environ.get("C", 512)), int(os.environ.get("IT", 600))
adapt = bool(int(os.environ.get("ADAPT", "0")))
pri = None
for K in [int(k) for k in os.environ.get("KS", "18,37,74").split(",")]:
    os.environ["GYNC_MCMC_SPEC"] = str(K)
    ss = [g.new_sampling(g.Config(g.Patient(la[i % 40], f"l{i % 40 + 1}"), thin=10, adapt=adapt), seed=7) for i in range(C_)]
    if pri is None:
        ss[0].config.prior(); pri = ss[0].config.priory0
    for s_ in ss: s_.config.priory0 = pri
    t0 = time.time(); g.sample_chains(ss, iters); dt = time.time() - t0
    print(f"C={C_} K={K:4d} adapt={adapt}: {dt:.2f} s  {C_ * iters / dt:.0f} chain-iterations/s  accept {np.mean([x.variate.naccept for x in ss]) / iters:.3f}", flush=True)
