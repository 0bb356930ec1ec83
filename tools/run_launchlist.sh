#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --quick --steps 3 --warmup 2 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:lg:: -s 16 -c 24 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
for r in rows[1:]:
    print(r[ix['Kernel Name']][:60].ljust(60), r[ix['Grid Size']].rjust(14), r[ix['Metric Value']].rjust(10), r[ix['Metric Unit']])
PY
