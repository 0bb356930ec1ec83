#!/bin/bash
# two GPUs: the rest of the GPU tests, then the album exchange across two ranks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','merged_album_check')})
for n,c in (d.get('configs') or {}).items(): print(n, {k:c[k] for k in ('value','ms_per_step','sweep_ms','frac')})
PY
