#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check | cut -c1-120
for xc in 0 1 2; do
echo "== XCLUSTER=$xc (0 = planner)"
LOUDGAIN_B200_XCLUSTER=$xc timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep quick | cut -c30-200
LOUDGAIN_B200_XCLUSTER=$xc LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --quick --steps 8 --warmup 3 2>&1 | grep "lgb step rank 0" | tail -1 | cut -c1-700
done
