from tail import *
z=np.load('/tmp/study/shim_1e-08_1e-10.npz')
tr=z['tr']
for b in [187,963,0,5]:
    e=np.abs(tr[b]-tro[b])/np.abs(tro[b])
    print(b,'rel traj err by time (max over h):'); print(np.array2string(e.max(1),precision=1,max_line_width=200))
    x=X[b]; print(' T=',x[115])
for b in [187,963]:
    for rt,at in [(1e-8,1e-10),(5e-9,1e-10),(3e-9,1e-10),(1e-9,1e-10),(1e-9,1e-11),(1e-8,1e-12),(1e-10,1e-12)]:
        st,t_,na_,nr_=shim_traj(np.ascontiguousarray(X[b]),rt,at)
        LL=np.array([o.llh_from_traj(t_,d,0.1) for d in datas])
        print(b,rt,at,np.abs(LL-LLo[b])/np.abs(LLo[b]),na_,nr_)
