#!/bin/bash
# Round-2 final pass on one GPU (second half of the round: pipelined runs): tests, the bench line,
# the reference arm, the ncu launch list and the full captures of the sweep and the kernels behind it.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > $o/r02_bench_n1.json 2> $o/r02_bench_n1.err; echo "bench rc=$?"; tail -c 300 $o/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $o/r02_bench_ref.json 2> $o/r02_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['truepeak_pass_ms'])
for n,c in (d.get('configs') or {}).items(): print(n, {k:c[k] for k in ('value','ms_per_step','sweep_ms','frac')})
r=json.load(open('gpurun_out/r02_bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])
PY
timeout 200 python bench.py --quick --steps 3 --warmup 2 > $o/r02_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 60 --csv \
    --log-file $o/r02_launch_list.csv \
    -k 'regex:run_sweep|tp_eval|fixslot|block_kernel|query_kernel|sweep_kernel|sweep_pair' \
    python bench.py --quick --steps 3 --warmup 2 > $o/r02_ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:run_sweep_kernel -s 2 -c 1 -o $o/r02_prof_s16 python bench.py --quick --steps 1 --warmup 1 > $o/r02_ncu_s16.log 2>&1
timeout 300 $NCU -k 'regex:tp_eval_run_kernel|fixslot_kernel|block_kernel|query_kernel' -s 10 -c 5 -o $o/r02_prof_post python bench.py --quick --steps 1 --warmup 1 > $o/r02_ncu_post.log 2>&1
for f in s16 post; do
  ncu -i $o/r02_prof_$f.ncu-rep --page raw --csv > $o/r02_prof_${f}_raw.csv 2>/dev/null
done
ls -la $o/r02_prof_*.ncu-rep $o/r02_prof_*_raw.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
