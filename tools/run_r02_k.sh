#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LOUDGAIN_B200_VERBOSE=1 timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | grep -E "quick|lgb\]"
bash tools/run_launchlist.sh 2>&1 | tail -9
