import numpy as np, time, torch
from tests.helpers import *
from oracle import load_oracle
from loudgain_b200 import synth
lib=load_oracle()
spec=synth.config1_spec(40.0)
t0=time.time(); pcm=synth.programme_s16(spec).numpy(); print("synth",time.time()-t0, pcm.shape, np.abs(pcm).max())
t0=time.time(); o=oracle_measure(lib,[(pcm,spec.rate)]); print("oracle",time.time()-t0)
t0=time.time(); e=emu_measure([(pcm,spec.rate)]); print("emu",time.time()-t0)
ot=o["tracks"][0]; et=e["tracks"][0]
print("oracle",ot["loudness"],ot["range"],ot["sample_peak"],ot["true_peak"],len(ot["blocks"]),len(ot["st"]))
print("emu   ",et["loudness"],et["range"],et["sample_peak"],et["true_peak"],et["n_abs"],et["n_st"],e["chunk_len"])
gate=10**((-70+0.691)/10)
eb=e["blocks"][e["blocks"]>=gate]; print("blocks rel err", rel_diff(eb,ot["blocks"]) if len(eb)==len(ot["blocks"]) else ("count mismatch",len(eb),len(ot["blocks"])))
es=e["st"][e["st"]>=gate]; print("st rel err", rel_diff(es,ot["st"]) if len(es)==len(ot["st"]) else "count mismatch")
print("dL",et["loudness"]-ot["loudness"],"dLRA",et["range"]-ot["range"],"tp rel",rel_diff(et["true_peak"],ot["true_peak"]))
