#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exchange" 2>&1 | tail -3
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check | cut -c1-200
timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep quick | cut -c30-330
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -2
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 800 gpurun_out/r02_bench_n2.err | grep -v "^\*\*\*\|OMP_NUM"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','merged_album_check')})
for n,c in (d.get('configs') or {}).items(): print(n, {k:c[k] for k in ('value','ms_per_step','sweep_ms','frac')})
PY
