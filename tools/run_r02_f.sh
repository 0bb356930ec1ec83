#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_f.txt
: > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
run() { echo "== variant=[$1] RUN_WARPS=$2" >> $out; LOUDGAIN_B200_VERBOSE=1 LG_LIB_SUFFIX=$1 LOUDGAIN_B200_RUN_WARPS=$2 timeout 300 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1; }
for v in A:8 A:10 A:11 B:12 B:16; do
  name=${v%%:*}; w=${v#*:}
  for x in "" nc nl nr nt; do run ${name}$x $w; done
done
python - <<'PY'
import re,json
cur=None
for line in open('gpurun_out/r02_f.txt'):
    if line.startswith('=='): cur=line.strip()
    elif line.startswith('[lgb]'): plan=re.search(r'L=(\d+) R=(\d+)',line).groups()
    elif line.startswith('{"quick"'):
        d=json.loads(line); print(cur, 'L,R=',plan, 'sweep %.4f tp %.4f step %.4f'%(d['sweep_ms'],d['truepeak_ms'],d['ms_per_step']))
PY
