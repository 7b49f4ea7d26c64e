"""A/B of the learned sample ordering inside the speculative Metropolis rounds (same process, same chains)."""
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
pri = None
res = {}
for name, env in (("sched", None), ("no-sched", "1"), ("sched", None)):
    if env: os.environ["GYNC_NO_SCHED"] = env
    else: os.environ.pop("GYNC_NO_SCHED", None)
    ss = [g.new_sampling(g.Config(g.Patient(la[i % 40], f"l{i % 40 + 1}"), thin=10), seed=7) for i in range(C_)]
    if pri is None:
        ss[0].config.prior(); pri = ss[0].config.priory0
    for s_ in ss: s_.config.priory0 = pri
    t0 = time.time(); g.sample_chains(ss, iters); dt = time.time() - t0
    res[name] = np.stack([s_.samples for s_ in ss])
    print(f"{name:9s} {C_} chains x {iters}: {dt:.2f} s  {C_ * iters / dt:.0f} chain-iterations/s", flush=True)
print("identical chains:", bool(np.array_equal(res["sched"], res["no-sched"])))
