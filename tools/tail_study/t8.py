from tail import *
rx=C.CDLL('/tmp/study/rx.so')
def rx_traj(x,rt,at,merge):
    ta=np.empty((62,4)); tb=np.empty((62,4)); na=C.c_int(); nb=C.c_int()
    st=rx.rx_traj(np.ascontiguousarray(x).ctypes.data_as(dp),C.c_double(rt),C.c_double(at),merge,ta.ctypes.data_as(dp),tb.ctypes.data_as(dp),C.byref(na),C.byref(nb))
    return st,ta,tb,na.value,nb.value
if __name__=="__main__":
 for rt,at,merge in [(1e-8,1e-10,2),(1e-7,1e-9,2),(1e-6,1e-8,2),(3e-7,3e-9,2)]:
     with ThreadPoolExecutor(8) as ex:
         res=list(ex.map(lambda i: rx_traj(X[i],rt,at,merge), range(NA)))
     LA=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
     LB=np.array([[o.llh_from_traj(r[2],d,0.1) for d in datas] for r in res])
     nA=np.mean([r[3] for r in res]); nB=np.mean([r[4] for r in res])
     EA=np.abs(LA-LLo)/np.abs(LLo); EB=np.abs(LB-LLo)/np.abs(LLo)
     print(rt,at,'merge',merge,'steps A %.0f B %.0f'%(nA,nB))
     print('  A: median %.2e p99 %.2e max %.2e'%(np.median(EA),np.quantile(EA,.99),EA.max()))
     print('  B: median %.2e p99 %.2e max %.2e ; ratio EB/EA quantiles'%(np.median(EB),np.quantile(EB,.99),EB.max()),np.quantile(EB/EA,[.01,.1,.5,.9,.99]))
     for k in [15,31,7]:
         # extrapolate on trajectories (then llh) 
         LX=np.array([[o.llh_from_traj(r[1]+(r[1]-r[2])/k,d,0.1) for d in datas] for r in res])
         EX=np.abs(LX-LLo)/np.abs(LLo)
         print('  extrap k=%d: median %.2e p99 %.2e max %.2e frac>1e-8 %.5f'%(k,np.median(EX),np.quantile(EX,.99),EX.max(),(EX>1e-8).mean()))
     D=np.abs(LA-LB)/np.abs(LLo)
     print('  est D/15 vs true EA: ratio EA/(D/15) quantiles',np.quantile(EA/(D/15),[0,.01,.5,.99,1]), 'logcorr',np.corrcoef(np.log(EA.ravel()),np.log(D.ravel()))[0,1])
