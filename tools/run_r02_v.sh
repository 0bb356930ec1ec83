#!/bin/bash
mkdir -p gpurun_out
for hold in 0 2; do
  echo "== TP_HOLD=$hold"
  for i in 1 2; do LOUDGAIN_B200_TP_HOLD=$hold timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c30-330; done
  LOUDGAIN_B200_TP_HOLD=$hold LOUDGAIN_B200_STEP_TRACE=1 timeout 300 python bench.py --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -1
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check | cut -c1-200
for hold in 0 2; do
  echo "== N=2 TP_HOLD=$hold"
  LOUDGAIN_B200_TP_HOLD=$hold timeout 300 $TR bench.py --gpus 2 --quick --steps 20 --warmup 3 2>&1 | grep quick | cut -c30-330
  LOUDGAIN_B200_TP_HOLD=$hold LOUDGAIN_B200_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --quick --steps 6 --warmup 3 2>&1 | grep "lgb step" | tail -2
done
