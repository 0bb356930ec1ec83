set -u
o=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err; tail -c 900 $o/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launch_list.csv -k 'regex:sweep_|tp_|truepeak_|fixup_|slot_|block_kernel|query_kernel' python bench.py --quick --steps 2 --warmup 1 > $o/ncu_launch.log 2>&1
wc -l $o/launch_list.csv
