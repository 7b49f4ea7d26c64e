"""Soak / determinism check of the persistent Metropolis kernel: the same chains run three times under different
scheduling (in-flight share 0.7, 1.1, 2.5 per lane; the last also on a shuffled chain order) must be bit-identical,
chain by chain; plus one long single chain against its own re-run.  python tests/manual/mcmc_soak.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gync_b200 as g
from gync_b200 import lib as L
L.check(L.load().gync_init(0))
la = g.api.alldatas()
c0 = g.Config(); c0.prior()
NC, IT = int(os.environ.get("NC", 1024)), int(os.environ.get("IT", 600))


def chains(order):
    ss = []
    for c in order:
        cf = g.Config(g.Patient(la[c % 40], f"p{c % 40}"), thin=5, propvar=g.api.uniformpropvar(0.1 if c % 3 else 0.01))
        cf.priory0 = c0.priory0
        ss.append(g.new_sampling(cf, seed=77, chain_id=c))
    return ss


ref = None
for over, shuffle in (("1.1", False), ("0.7", False), ("2.5", True)):
    os.environ["GYNC_MCMC_OVERSUB"] = over
    order = list(range(NC))
    if shuffle:
        np.random.default_rng(1).shuffle(order)
    ss = chains(order)
    t0 = time.time(); g.sample_chains(ss, IT); dt = time.time() - t0
    by_id = {s.variate.chain_id: s for s in ss}
    acc = np.mean([s.variate.naccept for s in ss]) / IT
    print(f"in-flight {over} per lane, shuffled={shuffle}: {NC * IT / dt:.0f} chain-iterations/s, accept {acc:.4f}", flush=True)
    if ref is None:
        ref = by_id
    else:
        bad = [c for c in range(NC) if not (np.array_equal(ref[c].samples, by_id[c].samples) and ref[c].variate.naccept == by_id[c].variate.naccept
                                             and ref[c].variate.logf == by_id[c].variate.logf)]
        assert not bad, bad[:10]
os.environ.pop("GYNC_MCMC_OVERSUB")
a = g.sample(g.Config(), 5000, seed=3)
b = g.sample_bang(g.sample(g.Config(), 2000, seed=3), 3000)
assert np.array_equal(a.samples, b.samples) and a.variate.naccept == b.variate.naccept
print("single chain 5000 == 2000 + 3000: ok, accept", a.variate.naccept / 5000)
print("soak ok")
