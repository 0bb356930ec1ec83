import numpy as np, scipy.signal as sg
from oracle import load_oracle, tp_phase
lib=load_oracle(); f32=np.float32
def test(factor, seed, kind):
    rng=np.random.default_rng(seed); n=400000
    if kind=="noise": x=rng.standard_normal(n); x/=np.abs(x).max()
    elif kind=="lp": x=sg.lfilter([1],[1,-0.9],rng.standard_normal(n)); x/=np.abs(x).max()
    else: x=np.clip(3*sg.lfilter([1],[1,-0.9],rng.standard_normal(n))/5,-1,1)
    pcm=np.clip(np.round(x*32767),-32768,32767).astype(np.int16)
    z=pcm.astype(np.float64)/32768  # exactly float-representable
    delay=(49+factor-1)//factor
    worst=0; tp_ref=0; tp_32=0
    for p in range(1,factor):
        c,s=tp_phase(lib,factor,p)
        # reference: acc double in tap order, out=(float)acc
        acc=np.zeros(n)
        zp=np.concatenate([np.zeros(delay),z])
        for ct,st in zip(c,s): acc+=zp[delay-st:delay-st+n]*ct
        ref=acc.astype(f32)
        # fp32 FMA chain on integer-valued floats, scaled at the end
        zi=np.concatenate([np.zeros(delay),pcm.astype(np.float64)])
        a32=np.zeros(n,f32); c32=c.astype(f32)
        for ct,st in zip(c32,s): a32=(zi[delay-st:delay-st+n]*np.float64(ct)+a32.astype(np.float64)).astype(f32)
        o32=(a32*f32(1/32768)).astype(f32)
        m=np.abs(ref).max(); tp_ref=max(tp_ref,m); tp_32=max(tp_32,np.abs(o32).max())
        top=np.abs(ref)>0.5*m
        worst=max(worst,(np.abs(o32[top].astype(np.float64)-ref[top])/np.abs(ref[top])).max())
    print(f"factor={factor} {kind}: tp_ref={tp_ref:.9f} tp_32={tp_32:.9f} rel={abs(tp_32-tp_ref)/tp_ref:.2e} worst-near-peak={worst:.2e}")
for f in (4,2):
    for k in ("noise","lp","clip"):
        test(f,3,k)
