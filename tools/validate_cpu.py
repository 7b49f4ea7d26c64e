#!/usr/bin/env python
"""Error distribution of the device solver's source (CPU build, tests/host_shim) at a given tolerance against the frozen
converged log-likelihoods (tests/golden/nearset_llh_golden.npz): sets A, B (18 432 near-set cases) and W (wide set).
    python tools/validate_cpu.py [rtol:atol ...]      -> JSON on stdout (committed as profiles/r2_validation_cpu.json)"""
import ctypes as C
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gync_oracle as o  # noqa: E402
import make_nearset_golden as mg  # noqa: E402

hs = C.CDLL(os.path.join(ROOT, "tests", "_build", "host_shim.so"))
dp = C.POINTER(C.c_double)


def traj(x, rt, at):
    tr = np.empty((62, 4))
    na, nr = C.c_int(), C.c_int()
    x = np.ascontiguousarray(x)
    st = hs.hs_meas_traj(x.ctypes.data_as(dp), C.c_double(rt), C.c_double(at), 200000, tr.ctypes.data_as(dp), C.byref(na), C.byref(nr))
    return st, tr, na.value + nr.value


if __name__ == "__main__":
    tols = [tuple(float(v) for v in a.split(":")) for a in sys.argv[1:]] or [(7e-10, 1e-11)]
    g = np.load(os.path.join(ROOT, "tests", "golden", "nearset_llh_golden.npz"))
    la = o.patients()[0]
    out = {"how": "python tools/validate_cpu.py " + " ".join(sys.argv[1:]) + "  (CPU build of csrc/gync_solver.cuh, Rodas5P)"}
    for rt, at in tols:
        key = f"rtol{rt:g}_atol{at:g}"
        out[key] = {}
        for name in ("A", "B", "C"):
            seed, n, pats = mg.SETS[name]
            X = mg.near_set(seed, n)
            with ThreadPoolExecutor(os.cpu_count()) as ex:
                res = list(ex.map(lambda i: traj(X[i], rt, at), range(n)))
            LL = np.array([[o.llh_from_traj(r[1], la[p], 0.1) for p in pats] for r in res])
            E = np.abs(LL - g["LL_" + name]) / np.abs(g["LL_" + name])
            out[key][name] = {"n_cases": int(E.size), "failures": int(sum(r[0] != 0 for r in res)),
                              "step_attempts_mean": float(np.mean([r[2] for r in res])),
                              "llh_rel_median": float(np.median(E)), "llh_rel_p99": float(np.quantile(E, 0.99)),
                              "llh_rel_max": float(E.max()), "frac_above_1e-8": float((E > 1e-8).mean())}
        Xw = mg.wide_set(int(g["seed_W"]), 512)
        with ThreadPoolExecutor(os.cpu_count()) as ex:
            res = list(ex.map(lambda i: traj(Xw[i], rt, at), range(512)))
        st = np.array([r[0] for r in res])
        fin = np.isfinite(g["LL_W"])
        LL = np.array([o.llh_from_traj(r[1], la[0], 0.1) if r[0] == 0 else -np.inf for r in res])
        both = fin & (st == 0)
        E = np.abs(LL[both] - g["LL_W"][both]) / np.abs(g["LL_W"][both])
        out[key]["W"] = {"n": 512, "fail_frac": float((st != 0).mean()), "oracle_fail_frac": float((~fin).mean()),
                         "failure_sets_equal": bool(np.array_equal(st == 0, fin)), "llh_rel_median": float(np.median(E)),
                         "llh_rel_max": float(E.max()), "step_attempts_mean_ok": float(np.mean([r[2] for r in res if r[0] == 0]))}
    print(json.dumps(out, indent=1))
