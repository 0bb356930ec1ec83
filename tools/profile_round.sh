#!/bin/bash
# One GPU-box pass for the numbers kept under profiles/: the bench line, the ncu
# launch list of the same command, and `ncu --set full` captures of the sweep
# (16-bit and float input) and of the post-sweep kernels.  Bench values come from
# the run WITHOUT a profiler.  Reports land in gpurun_out/; summarise them with
# tools/ncu_summary.py.
set -u
o=gpurun_out
python bench.py --steps 20 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err || exit 1
tail -c 600 $o/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launch_list.csv \
    -k 'regex:sweep_|tp_|truepeak_|fixup_|slot_|block_kernel|query_kernel' \
    python bench.py --quick --steps 2 --warmup 1 > $o/ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:sweep_pair_kernel -c 1 -o $o/prof_s16 python bench.py --quick --steps 1 --warmup 1 > $o/ncu_s16.log 2>&1
$NCU -k regex:sweep_pair_kernel -c 1 -o $o/prof_f32 python bench.py --quick --format f32 --steps 1 --warmup 1 > $o/ncu_f32.log 2>&1
$NCU -k 'regex:tp_scan_pair_kernel|tp_eval_pair_kernel|query_kernel' -c 3 -o $o/prof_post python bench.py --quick --steps 1 --warmup 1 > $o/ncu_post.log 2>&1
ls -la $o/*.ncu-rep
