"""A/B of the generic-grid forward solve (gync/forwardsol): tensor-memory kernel vs the older shared-memory kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gync_b200 as g
from gync_b200 import lib as L
L.check(L.load().gync_init(0))
N = int(os.environ.get("N", 18944))
rng = np.random.default_rng(4)
x0 = np.concatenate([g.api.refparms, g.api.refy0, [28.0]])
X = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((N, 116)))
res = {}
for name, env in (("tensor-memory", None), ("shared-memory", "1")):
    if env: os.environ["GYNC_SOLVE_SM"] = env
    g.forwardsol(X[:256])
    t0 = time.perf_counter(); Y = g.forwardsol(X); dt = time.perf_counter() - t0
    os.environ.pop("GYNC_SOLVE_SM", None)
    res[name] = Y
    print(f"forwardsol {N} samples x 31 times, {name} kernel: {dt*1e3:.1f} ms host call ({N/dt:.0f} solves/s incl. {Y.nbytes/1e6:.0f} MB D2H)")
a, b = res["tensor-memory"], res["shared-memory"]
ok = np.isfinite(a).all(axis=(1, 2)) & np.isfinite(b).all(axis=(1, 2))
print("finite in both:", int(ok.sum()), "of", N, " max rel diff:", float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300))))
