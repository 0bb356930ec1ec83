#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-420
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02_bench_n1c.json 2> gpurun_out/r02_bench_n1c.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1c.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')}, d['roofline']['kernel_ms'], d['roofline']['truepeak_pass_ms'])
PY
