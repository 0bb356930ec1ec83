#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/exchange_check.py 2>&1 | grep exchange_check | cut -c1-160
bash tools/run_bench_n.sh 2
