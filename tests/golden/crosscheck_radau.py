#!/usr/bin/env python
"""Cross-check EVERY finite sample of solution_golden.npz with an independent integrator.

The goldens are written by the C oracle's extrapolation integrator (make_solution_golden.py, which
checks 5 samples on the way).  This script re-solves all samples with scipy's Radau IIA (1e-12 /
1e-14, exact complex-step Jacobian) on the restated RHS and records, per sample, the largest relative
disagreement in log-likelihood (over the 4 patients) and in the measured trajectory.  The result is
committed as radau_crosscheck.json and asserted by tests/test_oracle.py.  Run in the build container:
    python tests/golden/crosscheck_radau.py        (~10 min on 8 cores)
"""
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)


def one(i):
    import gync_oracle as o
    from make_solution_golden import radau_traj
    g = np.load(os.path.join(HERE, "solution_golden.npz"))
    if not np.isfinite(g["LL"][i, 0]):
        return i, None, None, 0.0
    t0 = time.time()
    try:
        tr = radau_traj(g["X"][i])
    except AssertionError:
        return i, float("nan"), float("nan"), time.time() - t0
    rl = max(abs(o.llh_from_traj(tr, g["data"][:, :, m], float(g["sigma"])) - g["LL"][i, m]) / abs(g["LL"][i, m])
             for m in range(g["LL"].shape[1]))
    rt = float(np.max(np.abs(tr - g["traj"][i]) / np.abs(g["traj"][i])))
    return i, float(rl), rt, time.time() - t0


if __name__ == "__main__":
    g = np.load(os.path.join(HERE, "solution_golden.npz"))
    n = g["X"].shape[0]
    with Pool(int(os.environ.get("NPROC", "8"))) as pool:
        res = pool.map(one, range(n), chunksize=1)
    out = {"integrator": "scipy Radau IIA, rtol 1e-12, atol 1e-14, analytic (complex-step) Jacobian, restated RHS",
           "samples": [{"index": i, "llh_rel": rl, "traj_rel": rt, "seconds": round(s, 1)} for i, rl, rt, s in res]}
    fin = [r for r in res if r[1] is not None]
    out["n_checked"] = len(fin)
    out["n_nonfinite_golden"] = n - len(fin)
    out["max_llh_rel"] = max(r[1] for r in fin)
    out["max_traj_rel"] = max(r[2] for r in fin)
    with open(os.path.join(HERE, "radau_crosscheck.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "samples"}))
