import ctypes as C, os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gync_b200 as g
from gync_b200 import lib as L
lib=L.load(); L.check(lib.gync_init(0))
la=g.api.alldatas(); c0=g.Config(); c0.prior()
def chains(n):
    ss=[]
    for c in range(n):
        cf=g.Config(g.Patient(la[c%40],f"p{c%40}"),thin=10); cf.priory0=c0.priory0
        ss.append(g.new_sampling(cf,seed=1000,chain_id=c))
    return ss
g.sample_chains(chains(64),20)
for ov in ("1.1",):
    os.environ["GYNC_MCMC_OVERSUB"]=ov; ms=ov
    ss=chains(512); t0=time.time(); g.sample_chains(ss,2000); dt=time.time()-t0
    h=(C.c_uint64*24)(); f=C.c_uint64(); lib.gync_mcmc_last_histogram(h,C.byref(f))
    so,st,k=C.c_uint64(),C.c_uint64(),C.c_int32(); lib.gync_mcmc_last_stats(C.byref(so),C.byref(st),C.byref(k))
    print("max_steps",ms,"chain-it/s %.0f"%(512*2000/dt),"solves",so.value,"steps",st.value,"cancelled",lib.gync_mcmc_last_cancelled(),"early",lib.gync_mcmc_last_early(),"failed",f.value,"acc %.4f"%(np.mean([x.variate.naccept for x in ss])/2000))
    print("  hist log2(steps):",{b:int(h[b]) for b in range(24) if h[b]})
