#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --quick --steps 8 --warmup 3 2>&1 | grep "lgb step rank 0" | tail -3 | cut -c1-600
timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep quick | cut -c30-330
