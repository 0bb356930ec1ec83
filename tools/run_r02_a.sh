#!/bin/bash
# first GPU run of the run sweep: parity, then the quick bench at several residencies
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
for w in 8 12 16; do
  echo "== RUN_WARPS=$w"
  LOUDGAIN_B200_RUN_WARPS=$w timeout 300 python bench.py --quick --steps 20 --warmup 3 2>>gpurun_out/quick_err.log
done
echo "== legacy pair kernel"
LOUDGAIN_B200_RUN=0 timeout 300 python bench.py --quick --steps 20 --warmup 3 2>>gpurun_out/quick_err.log
