#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n1b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1b.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')})
PY
