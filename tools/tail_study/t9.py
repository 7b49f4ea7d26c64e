from tail import *
from t8 import rx_traj
rt,at=1e-8,1e-10
for merge in [2,3,4,6,8]:
    with ThreadPoolExecutor(8) as ex:
        res=list(ex.map(lambda i: rx_traj(X[i],rt,at,merge), range(NA)))
    LA=np.array([[o.llh_from_traj(r[1],d,0.1) for d in datas] for r in res])
    LB=np.array([[o.llh_from_traj(r[2],d,0.1) for d in datas] for r in res])
    bad=~np.isfinite(LB).all(1)
    EA=(np.abs(LA-LLo)/np.abs(LLo)).max(1)
    D=(np.abs(LA-LB)/np.abs(LA)).max(1); D[bad]=np.inf
    nB=np.mean([r[4] for r in res])
    print('merge',merge,'B steps %.0f'%nB,'nonfinite B',bad.sum(),'median D %.2e'%np.median(D))
    for lim in [1e-8,5e-9,3e-9]:
        tau=D[EA>lim].min()          # largest threshold that still flags every sample with EA>lim
        print('   catch all EA>%.0e: tau=%.2e flagged %.4f   (tau/2: flagged %.4f)'%(lim,tau,(D>=tau).mean(),(D>=tau/2).mean()))
