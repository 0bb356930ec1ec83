#!/bin/bash
mkdir -p gpurun_out
for hold in 1 0; do for qt in 1024 512; do
  echo "== TP_HOLD=$hold QUERY_THREADS=$qt"
  LOUDGAIN_B200_TP_HOLD=$hold LOUDGAIN_B200_QUERY_THREADS=$qt timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c30-330
  LOUDGAIN_B200_TP_HOLD=$hold LOUDGAIN_B200_QUERY_THREADS=$qt LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -1
done; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -3
