"""Steady-state many-chain Metropolis throughput (BASELINE configs[3] per-GPU share: 512 chains; and 4096 chains), long
enough that the last rounds — where only the unluckiest chains are still running — do not dominate."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o
L.check(L.load().gync_init(0))
la = o.patients()[0]
pri = None
for C_, iters, thin in ((512, int(os.environ.get("IT512", 2000)), 10), (4096, int(os.environ.get("IT4096", 1000)), 10)):
    ss = [g.new_sampling(g.Config(g.Patient(la[i % 40], f"l{i % 40 + 1}"), thin=thin), seed=7) for i in range(C_)]
    if pri is None:
        ss[0].config.prior(); pri = ss[0].config.priory0
    for s_ in ss: s_.config.priory0 = pri
    l0 = L.load().gync_launch_count()
    t0 = time.time(); g.sample_chains(ss, iters); dt = time.time() - t0
    acc = np.mean([x.variate.naccept for x in ss]) / iters
    print(f"{C_} chains x {iters} iterations (thin {thin}): {dt:.1f} s  {C_ * iters / dt:.0f} chain-iterations/s  accept {acc:.3f}  "
          f"launches {L.load().gync_launch_count() - l0}", flush=True)
