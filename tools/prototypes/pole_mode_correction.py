import numpy as np, scipy.signal as sg, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0, __import__("os").path.dirname(__file__))
from chunked_iir_fp32 import design, basis
f32=np.float32
def run(rate=44100,secs=10,k=10,W=72,bass=False,dc=0.0,seed=1):
    cf=design(rate); cf['q1']=cf['pb'][1]/cf['pb'][0]; cf['q2']=cf['pb'][2]/cf['pb'][0]
    s100=(rate+5)//10; L=s100//k
    c=cf['c']; e2=cf['e2']
    lam=np.roots([1,-(1+c),(c+e2)])[0]; 
    if lam.imag<0: lam=np.conj(lam)
    Hs=(1+cf['q1']/lam+cf['q2']/lam**2)/(1+cf['pa'][1]/lam+cf['pa'][2]/lam**2)
    x_,y_=(lam-1).real,(lam-1).imag
    def Aof(tau):  # tau=(d,w2) at f=0
        p=tau[1]/2; q=(tau[1]*x_-tau[0])/(2*y_); a=p+1j*q
        return 2*a*(lam-1)**2*Hs
    # check vs table
    N=W+L+12; al,be=basis(cf,N)
    f=np.arange(N)
    ea=np.real(Aof((1,0))*lam**f); eb=np.real(Aof((0,1))*lam**f)
    print("basis check f>=W: max|alpha-eig|/max|alpha| =",np.abs(al[W:]-ea[W:]).max()/np.abs(al[W:]).max(), np.abs(be[W:]-eb[W:]).max()/np.abs(be[W:]).max())
    # full chunked run with eigen accumulation in fp32
    rng=np.random.default_rng(seed); n=rate*secs; t=np.arange(n)/rate
    if bass: x=0.5*np.sin(2*np.pi*41*t)+0.01*rng.standard_normal(n)
    else:
        x=sg.lfilter([1],[1,-0.95],rng.standard_normal(n)); x/=np.abs(x).max(); x*=0.5*(1+0.8*np.sin(2*np.pi*0.13*t))
    x=x+dc
    pcm=np.clip(np.round(x*32767),-32768,32767).astype(np.int16)
    b=np.convolve(cf['pb'],[1,-2,1]); a=np.convolve(cf['pa'],cf['ra'])
    yref=sg.lfilter(b,a,pcm.astype(np.float64)/32768)
    nslots=n//s100; eref=(yref[:nslots*s100]**2).reshape(nslots,s100).sum(1)
    nch=nslots*k
    niter=(W+L+11)//12; NN=niter*12
    starts=np.arange(nch)*L-W
    xp=np.concatenate([np.zeros(W,dtype=np.int16),pcm,np.zeros(NN,dtype=np.int16)]).astype(f32)
    X=xp[starts[:,None]+W+np.arange(NN)[None,:]]
    e2f=f32(e2); cff=f32(c); p1=f32(cf['pa'][1]); p2=f32(cf['pa'][2]); q1=f32(cf['q1']); q2=f32(cf['q2'])
    lp=lam**np.arange(12); cr=lp.real.astype(f32); ci=lp.imag.astype(f32)
    rot=lam**-12; rr=f32(rot.real); ri=f32(rot.imag)
    z=np.zeros(nch,f32); d1=z.copy(); w1=z.copy(); w2=z.copy(); v1=z.copy(); v2=z.copy()
    E=np.zeros(nch,np.float64); Yr=z.copy(); Yi=z.copy()
    def fma(a,b,c_): return (a.astype(np.float64)*np.float64(b)+np.asarray(c_).astype(np.float64)).astype(f32)
    for it in range(niter):
        Sr=z.copy(); Si=z.copy(); Ep=z.copy()
        for i in range(12):
            f=it*12+i
            if f==W: P=(d1.copy(),w2.copy())
            xx=X[:,f]
            tt=fma(w2,-e2f,xx); d=fma(d1,cff,tt); w=w1+d; yh=d-d1
            u=fma(v2,-p2,yh); v=fma(v1,-p1,u); y=fma(v2,q2,fma(v1,q1,v))
            if W<=f<W+L:
                Ep=(y.astype(np.float64)*y+Ep).astype(f32); Sr=fma(y,cr[i],Sr); Si=fma(y,ci[i],Si)
            w2,w1,d1,v2,v1=w1,w,d,v1,v
            if f+1==W+L: Q=(d1.copy(),w2.copy())
        if it*12+12>W:
            E+=Ep
            nYr=fma(Yr,rr,fma(Yi,-ri,Sr)); nYi=fma(Yr,ri,fma(Yi,rr,Si)); Yr,Yi=nYr,nYi
    Xi=(Yr.astype(np.float64)+1j*Yi.astype(np.float64))*lam**((niter-1)*12)
    M=np.array([[c,-e2],[1.0,1.0]]); ML=np.linalg.matrix_power(M,L); MinvW=np.linalg.matrix_power(np.linalg.inv(M),W)
    S1=(np.abs(lam)**(2*np.arange(W,W+L))).sum(); S2=(lam**(2*np.arange(W,W+L))).sum()
    T=np.zeros(2); Et=np.zeros(nch)
    for j in range(nch):
        Pj=np.array([P[0][j],P[1][j]],dtype=np.float64); Qj=np.array([Q[0][j],Q[1][j]],dtype=np.float64)
        tau=MinvW@(T-Pj); A=Aof(tau)
        Et[j]=E[j]+2*np.real(A*Xi[j])+0.5*abs(A)**2*S1+0.5*np.real(A*A*S2)
        T=(Qj-ML@Pj)+ML@T
    g=(cf['pb'][0]/32768)**2
    eslot=Et.reshape(nslots,k).sum(1)*g
    rel=(eslot-eref)/eref
    print(f"rate={rate} k={k} W={W} bass={bass} dc={dc}: max|rel|={np.abs(rel).max():.3e} total rel={(eslot.sum()-eref.sum())/eref.sum():.3e} |Y|max={np.abs(Yr).max():.3g}")
run(); run(bass=True); run(dc=0.05); run(k=1,secs=6); run(rate=48000,W=84); run(rate=96000,W=156,secs=5); run(rate=192000,W=300,k=1,secs=3)
