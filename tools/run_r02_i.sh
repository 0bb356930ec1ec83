#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_i.txt
: > $out
run() { echo "== variant=[$1] RUN_WARPS=$2" >> $out; LOUDGAIN_B200_VERBOSE=1 LG_LIB_SUFFIX=$1 LOUDGAIN_B200_RUN_WARPS=$2 timeout 300 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1; }
run "" 16; run Bns 16; run Bnf 16; run Bnsf 16; run Bnt 16
python - <<'PY'
import re,json
cur=None
for line in open('gpurun_out/r02_i.txt'):
    if line.startswith('=='): cur=line.strip()
    elif line.startswith('[lgb]'): plan=re.search(r'L=(\d+) R=(\d+)',line).groups()
    elif line.startswith('{"quick"'):
        d=json.loads(line); print(cur, 'L,R=',plan, 'sweep %.4f tp %.4f step %.4f b2b %.4f cand %.3f'%(d['sweep_ms'],d['truepeak_ms'],d['ms_per_step'],d['ms_per_step_back_to_back'],d.get('tp_candidate_frac',-1)))
    elif 'Error' in line or 'error' in line: print(cur, line.strip()[:200])
PY
timeout 300 python bench.py --quick --steps 3 --warmup 2 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_sweep_kernel -s 4 -c 1 -o gpurun_out/r02_run_d python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu.log 2>&1
LG_LIB_SUFFIX=Bnt timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_sweep_kernel -s 4 -c 1 -o gpurun_out/r02_run_e python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu.log gpurun_out/ncu2.log
