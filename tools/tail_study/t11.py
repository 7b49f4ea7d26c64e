from tail import *
ph=C.CDLL('/tmp/study/ph.so')
def ph_traj(x,rt,at):
    tr=np.empty((62,4)); dt=np.empty((62,4))
    st=ph.ph_traj(np.ascontiguousarray(x).ctypes.data_as(dp),C.c_double(rt),C.c_double(at),tr.ctypes.data_as(dp),dt.ctypes.data_as(dp))
    return st,tr,dt
mag=np.array([120.,10.,400.,15.])*0.1
with ThreadPoolExecutor(8) as ex:
    res=list(ex.map(lambda i: ph_traj(X[i],1e-8,1e-10), range(NA)))
LL=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
E=np.abs(LL-LLo)/np.abs(LLo)
S=np.zeros((NA,3)); Sa=np.zeros((NA,3))
for i,r in enumerate(res):
    for m,d in enumerate(datas):
        tot=0; tota=0
        for p in range(2):
            g=(d-r[1][31*p:31*p+31])/mag**2*r[2][31*p:31*p+31]
            msk=~np.isnan(d); tot+=g[msk].sum(); tota+=np.abs(g[msk]).sum()
        S[i,m]=abs(tot)/abs(LL[i,m]); Sa[i,m]=tota/abs(LL[i,m])
for name,s in [('signed',S),('abs',Sa)]:
    print(name,'logcorr',np.corrcoef(np.log(E.ravel()),np.log(s.ravel()))[0,1],'ratio E/S quantiles',np.quantile(E/s,[0,.01,.5,.99,1]))
    rk=np.argsort(np.argsort(-s.max(1)))
    Em=E.max(1)
    for lim in [1e-8,5e-9,3e-9]: print('   flag frac to catch all E>%g: %.4f'%(lim,(rk[Em>lim].max()+1)/NA))
