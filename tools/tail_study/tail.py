import os, sys, time, ctypes as C
import numpy as np
from concurrent.futures import ThreadPoolExecutor
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/oracle')
import gync_oracle as o, c_oracle as co
hs=C.CDLL(ROOT+'/tests/_build/host_shim.so')
dp=C.POINTER(C.c_double)
def shim_traj(x, rt, at):
    tr=np.empty((62,4)); na=C.c_int(); nr=C.c_int()
    st=hs.hs_meas_traj(x.ctypes.data_as(dp), C.c_double(rt), C.c_double(at), 200000, tr.ctypes.data_as(dp), C.byref(na), C.byref(nr))
    return st, tr, na.value, nr.value
NA=int(os.environ.get('NA',2048))
rng=np.random.Generator(np.random.PCG64(99))
X=np.exp(np.log(o.init_x())[None,:]+0.0998*rng.standard_normal((NA,116)))
la=o.patients()[0]
datas=[la[0],la[1],la[4]]
cache='/tmp/study/oracle_%d.npz'%NA
if os.path.exists(cache):
    z=np.load(cache); LLo=z['LL']; tro=z['tr']
else:
    t0=time.time(); LLo,tro=co.llh_batch(X,datas,[0.1]*3,1,1e-12,1e-12,want_traj=True); print('oracle',time.time()-t0)
    np.savez(cache,LL=LLo,tr=tro)
def run(rt,at):
    with ThreadPoolExecutor(8) as ex:
        res=list(ex.map(lambda i: shim_traj(np.ascontiguousarray(X[i]),rt,at), range(NA)))
    LL=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
    return LL, np.array([r[2] for r in res]), np.array([r[3] for r in res]), np.array([r[1] for r in res])
if __name__=='__main__':
    for rt,at in [(1e-8,1e-10)]:
        t0=time.time(); LL,na,nr,tr=run(rt,at); print('shim',time.time()-t0)
        E=np.abs(LL-LLo)/np.abs(LLo)
        np.savez('/tmp/study/shim_%g_%g.npz'%(rt,at),LL=LL,na=na,nr=nr,tr=tr)
        print(rt,at,'median',np.median(E),'p99',np.quantile(E,.99),'max',E.max(),'frac',(E>1e-8).mean(), 'steps',na.mean(),nr.mean())
        bad=np.where(E.max(1)>5e-9)[0]
        for b in bad: print(b,E[b],na[b],nr[b])
