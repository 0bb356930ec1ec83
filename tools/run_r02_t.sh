#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-420
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -2
LOUDGAIN_B200_TPEVAL_CTAS=8 timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-420
timeout 1500 bash tools/profile_r02.sh 2>&1 | tail -12
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launch_list.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
for r in rows[1:][-22:]:
    print(r[ix['Kernel Name']][:70].ljust(70), r[ix['Grid Size']].rjust(14), r[ix['Block Size']].rjust(14), r[ix['Metric Value']].rjust(10))
PY
