#!/bin/bash
# one GPU: parity after the candidate-queue change, then step traces by query CTA size / tp_eval cap
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for qt in 1024 512 256; do for tc in 6 4; do
  echo "== QUERY_THREADS=$qt TPEVAL_CTAS=$tc"
  LOUDGAIN_B200_QUERY_THREADS=$qt LOUDGAIN_B200_TPEVAL_CTAS=$tc timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-420
  LOUDGAIN_B200_QUERY_THREADS=$qt LOUDGAIN_B200_TPEVAL_CTAS=$tc LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -2
done; done
