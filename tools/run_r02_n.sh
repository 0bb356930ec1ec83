#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/e2e_sweep.py 2>&1 | tail -6
LOUDGAIN_B200_NT_MIN=65536 timeout 300 python tools/e2e_sweep.py 2>&1 | tail -6
LOUDGAIN_B200_NT_MIN=100000000 timeout 300 python tools/e2e_sweep.py 2>&1 | tail -6
