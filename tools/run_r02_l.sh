#!/bin/bash
# single GPU: new GPU tests, the exchange check with one rank, a quick bench and the full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/exchange_check.py 2>&1 | tail -2
timeout 300 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
