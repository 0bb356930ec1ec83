#!/bin/bash
# Round-2 profile pass (one GPU): the ncu launch list of the bench command and `ncu --set full`
# captures of the run sweep (16-bit and float input) and of the kernels behind it.  Bench values
# come from runs WITHOUT a profiler (tools/run_r02_*.sh); summarise with tools/ncu_summary.py.
set -u
o=gpurun_out
python bench.py --quick --steps 3 --warmup 2 > $o/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 60 --csv \
    --log-file $o/r02_launch_list.csv \
    -k 'regex:run_sweep|tp_filter|tp_eval|fixslot|block_kernel|query_kernel|sweep_kernel|sweep_pair' \
    python bench.py --quick --steps 3 --warmup 2 > $o/r02_ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:run_sweep_kernel -s 2 -c 1 -o $o/r02_prof_s16 python bench.py --quick --steps 1 --warmup 1 > $o/r02_ncu_s16.log 2>&1
$NCU -k regex:run_sweep_kernel -s 2 -c 1 -o $o/r02_prof_f32 python bench.py --quick --format f32 --steps 1 --warmup 1 > $o/r02_ncu_f32.log 2>&1
$NCU -k 'regex:tp_eval_run_kernel|fixslot_kernel|block_kernel|query_kernel' -s 10 -c 5 -o $o/r02_prof_post python bench.py --quick --steps 1 --warmup 1 > $o/r02_ncu_post.log 2>&1
for f in s16 f32 post; do
  ncu -i $o/r02_prof_$f.ncu-rep --page raw --csv > $o/r02_prof_${f}_raw.csv 2>/dev/null
done
ls -la $o/r02_prof_*.ncu-rep $o/r02_prof_*_raw.csv
