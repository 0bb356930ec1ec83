#!/bin/bash
for i in 1 2 3; do timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c30-330; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
