#!/bin/bash
# Tuning sweep of the packed sweep kernel on the GPU box: library variants
# (built with LG_LIB_SUFFIX / LG_NVCC_EXTRA) x chunks per 100 ms slot.
out=gpurun_out/tune_pair.txt
: > $out
for v in "" ca ca128 r3 r4; do
  for k in 5 7; do
    echo "variant=[$v] k=$k" >> $out
    LG_LIB_SUFFIX=$v LOUDGAIN_B200_CHUNKS_PER_SLOT=$k timeout 120 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1
  done
done
cat $out
