from tail import *
z=np.load('/tmp/study/shim_1e-08_1e-10.npz'); LLf=z['LL']; naf=z['na']; nrf=z['nr']
Ef=(np.abs(LLf-LLo)/np.abs(LLo)).max(1)
worst=np.argsort(-Ef)
print('fine steps corr with log Ef', np.corrcoef(naf,np.log(Ef))[0,1], 'nrej',np.corrcoef(nrf,np.log(Ef))[0,1])
print('rank by steps of worst 10', [int((naf>naf[i]).sum()) for i in worst[:10]])
for rt,at in [(1e-4,1e-6),(1e-5,1e-7),(3e-6,3e-8)]:
    LL,na,nr,tr=run(rt,at)
    D=(np.abs(LL-LLf)/np.abs(LLf)).max(1)
    rk=np.argsort(np.argsort(-D))
    n5=(Ef>5e-9).sum(); n3=(Ef>3e-9).sum()
    print(rt,at,'steps',(na+nr).mean(),'flag frac needed to catch all Ef>5e-9 (%d): %.4f ; all Ef>3e-9 (%d): %.4f ; >1e-8: %.4f'%(n5,(rk[Ef>5e-9].max()+1)/NA,n3,(rk[Ef>3e-9].max()+1)/NA,(rk[Ef>1e-8].max()+1)/NA))
