#!/bin/bash
# Round-2 final pass on one GPU: tests, the bench line, the reference arm, the ncu profile pass.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['truepeak_pass_ms'])
for n,c in (d.get('configs') or {}).items(): print(n, {k:c[k] for k in ('value','ms_per_step','sweep_ms','frac')})
r=json.load(open('gpurun_out/r02_bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])
PY
timeout 1500 bash tools/profile_r02.sh 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
