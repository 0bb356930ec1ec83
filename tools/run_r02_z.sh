#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check | cut -c1-200
for i in 1 2; do timeout 300 $TR bench.py --gpus $N --quick --steps 20 --warmup 3 2>&1 | grep quick | cut -c30-330; done
LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus $N --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -1 | cut -c1-400
