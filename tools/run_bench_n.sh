#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it; the line lands in gpurun_out/r02_bench_n$N.json
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612"; fi
timeout 1200 $L bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|^W1019\|^$" gpurun_out/r02_bench_n$N.err | tail -15
python - "$N" <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/r02_bench_n{sys.argv[1]}.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','merged_album_check','gpu_launches')})
for n,c in (d.get('configs') or {}).items(): print(n, {k:c[k] for k in ('value','ms_per_step','sweep_ms','frac')})
PY
