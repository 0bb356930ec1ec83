#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
LOUDGAIN_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"; grep -a "lgb_scan_host" gpurun_out/r02_bench_n1b.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1b.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')})
PY
LOUDGAIN_B200_TRACE=1 timeout 300 python tools/e2e_sweep.py 2>&1 | tail -30
