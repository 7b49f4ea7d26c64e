from tail import *
for rt,at in [(1e-8,1e-12),(5e-9,1e-12),(3e-9,1e-12),(2e-9,1e-12),(2e-9,1e-11),(1e-9,1e-11)]:
    LL,na,nr,tr=run(rt,at)
    E=np.abs(LL-LLo)/np.abs(LLo)
    print(rt,at,'median %.2e p99 %.2e max %.2e frac %.5f steps %.0f'%(np.median(E),np.quantile(E,.99),E.max(),(E>1e-8).mean(),(na+nr).mean()), 'worst',np.argsort(-E.max(1))[:4], np.sort(E.max(1))[-4:])
