#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_b.txt
: > $out
for v in "" nl nc; do
  for w in 8 16; do
    echo "== variant=[$v] RUN_WARPS=$w" >> $out
    LOUDGAIN_B200_VERBOSE=1 LG_LIB_SUFFIX=$v LOUDGAIN_B200_RUN_WARPS=$w timeout 300 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1
  done
done
grep -E "==|quick|lgb\]" $out | sort -u | head -40
timeout 300 python bench.py --quick --steps 3 --warmup 2 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_sweep_kernel -s 4 -c 1 -o gpurun_out/r02_run_a python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
