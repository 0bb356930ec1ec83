#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_c.txt
: > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
run() { echo "== variant=[$1] RUN_WARPS=$2" >> $out; LOUDGAIN_B200_VERBOSE=1 LG_LIB_SUFFIX=$1 LOUDGAIN_B200_RUN_WARPS=$2 timeout 300 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1; }
run "" 8; run "" 11
run r2 8; run r2 12; run r2 16
run s24 8; run s24 6
run w8 8
run nl 8; run nl 16; run nc 8; run nc 16
grep -E "==|quick|lgb\]" $out | cut -c1-250
