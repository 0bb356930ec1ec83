#!/bin/bash
# Tuning sweep of the packed sweep kernel on the GPU box: library variants
# (built with LG_LIB_SUFFIX / LG_NVCC_EXTRA), 16-bit and float input, TMA staging on.
out=gpurun_out/tune_pair.txt
: > $out
for v in "" c d e f; do
  for fmt in s16 f32; do
    echo "variant=[$v] fmt=$fmt" >> $out
    LOUDGAIN_B200_TMA=1 LG_LIB_SUFFIX=$v timeout 120 python bench.py --quick --format $fmt --steps 10 --warmup 3 >> $out 2>&1
  done
done
cat $out
