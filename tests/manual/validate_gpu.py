"""Broader validation + secondary configs on a GPU box (not part of the test-suite: ~3 min).
 (A) log-likelihood error distribution of the CUDA path vs the C oracle (1e-12) on a near-set sample x 3 patients
 (B) the 'wide' set (eb/gync.jl:59): failure fraction, agreement of the failure SET and of the finite values
 (C) BASELINE configs[0]: sample(Config(), 1000) single chain;  configs[3] scaled: 4096 chains x 20 iterations"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o, c_oracle as co
L.check(L.load().gync_init(0))
out = {}
NA = int(os.environ.get("NA", 1024))
rng = np.random.Generator(np.random.PCG64(99))
X = np.exp(np.log(o.init_x())[None, :] + 0.0998 * rng.standard_normal((NA, 116)))
la = o.patients()[0]
datas = [la[0], la[1], la[4]]
t0 = time.time(); LLo, tro = co.llh_batch(X, datas, [0.1] * 3, 1, 1e-12, 1e-12, want_traj=True); t_or = time.time() - t0
TOLS = [tuple(float(v) for v in t.split(":")) for t in os.environ.get("TOLS", "1e-8:1e-10,1e-8:1e-8").split(",")]
for rt, at in TOLS:
    LL, st, Ym = g.loglik_matrix(X, datas, [0.1] * 3, rt, at, want_traj=True)
    ok = (st == 0) & np.isfinite(LLo[:, 0])
    E = np.abs(LL[ok] - LLo[ok]) / np.abs(LLo[ok])
    Et = np.max(np.abs(Ym[ok] - tro[ok]) / np.abs(tro[ok]), axis=(1, 2))
    out[f"A_rtol{rt:g}_atol{at:g}"] = {"n_cases": int(E.size), "llh_rel_median": float(np.median(E)), "llh_rel_p99": float(np.quantile(E, 0.99)),
                                   "llh_rel_max": float(E.max()), "frac_above_1e-8": float((E > 1e-8).mean()),
                                   "traj_rel_max": float(Et.max()), "failures_gpu": int((st != 0).sum())}
out["A_oracle_seconds"] = t_or
if os.environ.get("ONLY_A"):
    print(json.dumps(out, indent=1)); sys.exit(0)
# (B) wide set
gold = np.load(os.path.join(ROOT, "tests", "golden", "solution_golden.npz"))
means, stds = gold["gmm_means"], gold["gmm_stds"]
NW = 512
Xw = np.array([np.concatenate([o.REFPARMS * rng.random(82) * 5, means[rng.integers(31)] + stds * rng.standard_normal(33), [30.0]]) for _ in range(NW)])
LLw, stw, nsw = g.loglik_matrix(Xw, [la[0]], [0.1], want_steps=True)
LLwo = co.llh_batch(Xw, [la[0]], [0.1], 1, 1e-11, 1e-11)[:, 0]
both = np.isfinite(LLw[:, 0]) & np.isfinite(LLwo)
Ew = np.abs(LLw[both, 0] - LLwo[both]) / np.abs(LLwo[both])
out["B_wide"] = {"n": NW, "gpu_fail_frac": float((stw != 0).mean()), "oracle_fail_frac": float((~np.isfinite(LLwo)).mean()),
                 "failure_sets_equal": bool(np.array_equal(stw != 0, ~np.isfinite(LLwo))), "n_disagree": int(((stw != 0) != ~np.isfinite(LLwo)).sum()),
                 "llh_rel_median": float(np.median(Ew)) if Ew.size else None, "llh_rel_max": float(Ew.max()) if Ew.size else None,
                 "mean_steps_ok": float(nsw[stw == 0].sum(1).mean()) if (stw == 0).any() else None}
# (C) chains
c = g.Config()
t0 = time.time(); s = g.sample(c, 1000, seed=1); dt = time.time() - t0
out["C_cfg0_single_chain_1000"] = {"seconds": dt, "iters_per_s": 1000 / dt, "accept_rate": s.variate.naccept / 1000, "shape": list(s.samples.shape)}
pats = [g.Patient(la[i % 40], f"l{i % 40 + 1}") for i in range(4096)]
ss = [g.new_sampling(g.Config(p), seed=7) for p in pats]
ss[0].config.prior()
pri = ss[0].config.priory0
for s_ in ss: s_.config.priory0 = pri
t0 = time.time(); g.sample_chains(ss, 20); dt = time.time() - t0
out["C_cfg3_4096_chains_x20"] = {"seconds": dt, "chain_iters_per_s": 4096 * 20 / dt, "mean_accept": float(np.mean([x.variate.naccept for x in ss]) / 20)}
print(json.dumps(out, indent=1))
