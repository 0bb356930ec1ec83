#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check
timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep quick
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -4
