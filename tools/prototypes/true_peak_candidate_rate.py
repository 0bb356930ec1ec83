import numpy as np, sys
sys.path.insert(0,'/root/repo')
from loudgain_b200 import synth
from oracle import load_oracle
lib=load_oracle()
G=1.8642
def rate(pcm, tp):
    x=np.abs(pcm.astype(np.float32))/32768.0
    n=(len(x)//12)*12
    m12=x[:n].reshape(-1,12).max(1)
    mprev=np.concatenate([[0],m12[:-1]])
    cand=(np.maximum(m12,mprev)*G>tp)
    return cand.mean()
specs=synth.config2_specs(12, scale=0.25)
for s in specs[:6]:
    pcm=synth.programme_s16(s).numpy()
    st=lib.init(2,s.rate); st.add_frames(pcm, 65536); tps=st.true_peaks(); sps=st.sample_peaks(); st.destroy()
    print(f"level {s.level_db:+.1f} dB  tp={tps[0]:.3f}/{tps[1]:.3f} sp={sps[0]:.3f}  cand vs TP: {rate(pcm[:,0],tps[0]):.3f} {rate(pcm[:,1],tps[1]):.3f}   vs SP: {rate(pcm[:,0],sps[0]):.3f}")
# white noise and a 'mastered' (clipped) signal for contrast
rng=np.random.default_rng(1)
w=(rng.standard_normal(44100*30)*0.25); w=np.clip(w,-1,1); pcm=(w*32767).astype(np.int16)
print("white gaussian:", rate(pcm, np.abs(w).max()))
c=np.clip(rng.standard_normal(44100*30)*0.6,-0.98,0.98); pcm=(c*32767).astype(np.int16)
print("hard-clipped noise:", rate(pcm, 1.3))
