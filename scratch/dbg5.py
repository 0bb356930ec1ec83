import numpy as np
from loudgain_b200 import synth
from tests.helpers import *
from oracle import load_oracle
lib=load_oracle()
spec = synth.TrackSpec(seed=48000+5, rate=48000, channels=5, seconds=8.0, surround_db=1.5)
pcm=synth.programme_s16(spec).numpy()
o=oracle_measure(lib,[(pcm,48000)])["tracks"][0]; e=emu_measure([(pcm,48000)])["tracks"][0]
print(o["true_peak"], e["true_peak"])
# single-channel check of channel 1
p1=np.ascontiguousarray(pcm[:,1:2]); o1=oracle_measure(lib,[(p1,48000)])["tracks"][0]; e1=emu_measure([(p1,48000)])["tracks"][0]
print(o1["true_peak"], e1["true_peak"])
# where is the peak? brute-force FIR
from oracle import tp_phase
x=p1[:,0].astype(np.float64)/32768
best=(0,0,0)
for ph in (1,2,3):
    c,s=tp_phase(lib,4,ph)
    y=np.convolve(x,c)[:len(x)]
    i=np.argmax(np.abs(y)); 
    if abs(y[i])>best[0]: best=(abs(y[i]),ph,i)
print(best, len(x), e["n_abs"])
