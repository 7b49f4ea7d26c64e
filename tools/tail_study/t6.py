from tail import *
av=C.CDLL('/tmp/study/av.so')
def av_traj(x,rt,atv,norm=0):
    tr=np.empty((62,4)); ns=C.c_int(); ymin=np.empty(33); ymax=np.empty(33)
    atv=np.ascontiguousarray(atv,dtype=np.float64)
    st=av.av_traj(np.ascontiguousarray(x).ctypes.data_as(dp),C.c_double(rt),atv.ctypes.data_as(dp),norm,tr.ctypes.data_as(dp),C.byref(ns),ymin.ctypes.data_as(dp),ymax.ctypes.data_as(dp))
    return st,tr,ns.value,ymin,ymax
def err(b,tr): 
    LL=np.array([o.llh_from_traj(tr,d,0.1) for d in datas]); return (np.abs(LL-LLo[b])/np.abs(LLo[b])).max()
if __name__=='__main__':
    for b in [187,963]:
        st,tr,ns,ymin,ymax=av_traj(X[b],1e-8,np.full(33,1e-10))
        print(b,'base',err(b,tr),ns)
        print(' ymin',np.array2string(ymin,precision=1,max_line_width=250)); print(' ymax',np.array2string(ymax,precision=1,max_line_width=250))
        for i in range(33):
            a=np.full(33,1e-10); a[i]=1e-14
            st,tr,ns,_,_=av_traj(X[b],1e-8,a)
            print('  comp',i,'err %.2e steps %d'%(err(b,tr),ns))
        st,tr,ns,_,_=av_traj(X[b],1e-8,np.full(33,1e-10),1); print(b,'maxnorm',err(b,tr),ns)
