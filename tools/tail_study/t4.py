from tail import *
z=np.load('/tmp/study/shim_1e-08_1e-10.npz'); LLf=z['LL']; naf=z['na']
Ef=(np.abs(LLf-LLo)/np.abs(LLo)).max(1)
for rt,at in [(1e-5,1e-7),(1e-6,1e-8),(1e-6,1e-10)]:
    LL,na,nr,tr=run(rt,at)
    Ec=(np.abs(LL-LLo)/np.abs(LLo)).max(1)       # true coarse error
    D=(np.abs(LL-LLf)/np.abs(LLf)).max(1)        # observable: coarse-fine disagreement
    print(rt,at,'steps %.0f (%.2f of fine)'%((na+nr).mean(),(na+nr).mean()/5022),'coarse err median %.2e'%np.median(Ec))
    print(' log-corr(D,Ef)=%.3f'%np.corrcoef(np.log(D),np.log(Ef))[0,1])
    order=np.argsort(-D)
    rank_of={i:r for r,i in enumerate(order)}
    worst=np.argsort(-Ef)[:12]
    print(' rank in D of the 12 worst fine-error samples:',[rank_of[i] for i in worst])
    print(' ratio Ef/D quantiles',np.quantile(Ef/D,[.01,.5,.99,1.0]))
