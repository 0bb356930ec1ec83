#!/bin/bash
mkdir -p gpurun_out
export LG_LIB_SUFFIX=A LOUDGAIN_B200_RUN_WARPS=8
timeout 300 python bench.py --quick --steps 3 --warmup 2 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_sweep_kernel -s 4 -c 1 -o gpurun_out/r02_run_b python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
export LG_LIB_SUFFIX=B LOUDGAIN_B200_RUN_WARPS=16
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_sweep_kernel -s 4 -c 1 -o gpurun_out/r02_run_c python bench.py --quick --steps 3 --warmup 2 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
