from tail import *
es=C.CDLL('/tmp/study/eps.so')
def eps_traj(x,rt,at,variant):
    tr=np.empty((62,4)); ep=np.empty((62,4)); na=C.c_int()
    st=es.eps_traj(x.ctypes.data_as(dp),C.c_double(rt),C.c_double(at),variant,tr.ctypes.data_as(dp),ep.ctypes.data_as(dp),C.byref(na))
    return st,tr,ep,na.value
mag=np.array([120.,10.,400.,15.])*0.1
def dll(tr,ep,d):
    # linearised change of llh: sum r*isc^2*eps over observed entries, both periods
    tot=0.0; tota=0.0
    for p in range(2):
        r=(d-tr[31*p:31*p+31])/mag; g=r/mag*ep[31*p:31*p+31]
        m=~np.isnan(d)
        tot+=g[m].sum(); tota+=np.abs(g[m]).sum()
    return tot,tota
rt,at=1e-8,1e-10
import sys
for variant in [2,3,0]:
    with ThreadPoolExecutor(8) as ex:
        res=list(ex.map(lambda i: eps_traj(np.ascontiguousarray(X[i]),rt,at,variant), range(NA)))
    LL=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
    Et=np.abs(LL-LLo)      # true abs error
    est=np.array([[dll(r[1],r[2],d) for d in datas] for r in res])   # (NA,3,2)
    for name,e in [('signed',np.abs(est[:,:,0])),('abs',est[:,:,1])]:
        ratio=Et/e
        print('variant',variant,name,'log-corr %.3f'%np.corrcoef(np.log(Et.ravel()),np.log(e.ravel()))[0,1],'ratio true/est quantiles(.001,.01,.5,.99,1):',np.quantile(ratio,[.001,.01,.5,.99,1.0]))
        # flag rule: est*c > 1e-8*|LL| ; choose c = max ratio*1.0 ; count flagged
        relest=e/np.abs(LLo)
        for c in [np.quantile(ratio,1.0),0.1,0.05,0.03]:
            flagged=(relest*c>1e-8).any(1)
            missed=((Et/np.abs(LLo)>1e-8).any(1)&~flagged).sum()
            print('   c=%.3g flagged %.4f missed %d'%(c,flagged.mean(),missed))
