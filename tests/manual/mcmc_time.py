import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o
L.check(L.load().gync_init(0))
la = o.patients()[0]
c = g.Config()
t0 = time.time(); s = g.sample(c, 1000, seed=1); dt = time.time() - t0
print("cfg0 single chain 1000 iters: %.2f s  (%.1f iters/s) accept %.3f" % (dt, 1000/dt, s.variate.naccept/1000))
pats = [g.Patient(la[i % 40], f"l{i % 40 + 1}") for i in range(4096)]
ss = [g.new_sampling(g.Config(p), seed=7) for p in pats]
ss[0].config.prior(); pri = ss[0].config.priory0
for s_ in ss: s_.config.priory0 = pri
t0 = time.time(); g.sample_chains(ss, 100); dt = time.time() - t0
print("cfg3 4096 chains x 100: %.2f s  (%.0f chain-iters/s) accept %.3f" % (dt, 4096*100/dt, np.mean([x.variate.naccept for x in ss])/100))
# Config.adapt = true (Mamba's adaptation; the mode the reference's production notebooks sample with)
ca = g.Config(adapt=True)
t0 = time.time(); s = g.sample(ca, 1000, seed=1); dt = time.time() - t0
print("cfg0 adaptive single chain 1000 iters: %.2f s  (%.1f iters/s) accept %.3f  trace(propadapt) %.4g" % (
    dt, 1000/dt, s.variate.naccept/1000, np.trace(g.api.propadapt(s))))
ss = [g.new_sampling(g.Config(p, adapt=True), seed=7) for p in pats[:1024]]
for s_ in ss: s_.config.priory0 = pri
t0 = time.time(); g.sample_chains(ss, 50); dt = time.time() - t0
print("adaptive 1024 chains x 50: %.2f s  (%.0f chain-iters/s) accept %.3f" % (dt, 1024*50/dt, np.mean([x.variate.naccept for x in ss])/50))
