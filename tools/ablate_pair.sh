#!/bin/bash
# Ablations of the packed sweep on the GPU box: no-load / no-compute variants,
# then an ncu capture of the product kernel.
out=gpurun_out/ablate_pair.txt
: > $out
for v in "" nl nc c cnl cnc; do
  echo "variant=[$v] k=5" >> $out
  LG_LIB_SUFFIX=$v LOUDGAIN_B200_CHUNKS_PER_SLOT=5 timeout 120 python bench.py --quick --steps 10 --warmup 3 >> $out 2>&1
done
cat $out
LOUDGAIN_B200_CHUNKS_PER_SLOT=5 ncu --set full --clock-control none --import-source on -k regex:"sweep_pair_kernel|truepeak_pair_kernel" -c 2 -o gpurun_out/prof_pair1 python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu_pair1.log 2>&1
