from tail import *
ct=C.CDLL('/tmp/study/ct.so')
def ct_traj(x,rt,at,safe,fac1,expo,gus):
    tr=np.empty((62,4)); ns=C.c_int()
    st=ct.ct_traj(np.ascontiguousarray(x).ctypes.data_as(dp),C.c_double(rt),C.c_double(at),C.c_double(safe),C.c_double(fac1),C.c_double(expo),gus,tr.ctypes.data_as(dp),C.byref(ns))
    return st,tr,ns.value
for (rt,at,safe,fac1,expo,gus) in [(1e-8,1e-10,.9,5,.25,1),(1e-8,1e-10,.9,2,.25,1),(1e-8,1e-10,.9,1.3,.25,1),(1e-8,1e-10,.9,5,.2,1),(1e-8,1e-10,.9,5,.25,0),(2e-8,2e-10,.9,1.3,.25,1),(1e-8,1e-10,.8,1.5,.25,1)]:
    with ThreadPoolExecutor(8) as ex:
        res=list(ex.map(lambda i: ct_traj(X[i],rt,at,safe,fac1,expo,gus), range(NA)))
    LL=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
    E=np.abs(LL-LLo)/np.abs(LLo)
    print((rt,at,safe,fac1,expo,gus),'steps %.0f median %.2e p99 %.2e max %.2e frac %.5f'%(np.mean([r[2] for r in res]),np.median(E),np.quantile(E,.99),E.max(),(E>1e-8).mean()))
