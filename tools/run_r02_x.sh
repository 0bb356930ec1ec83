#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c30-330; done
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -1
