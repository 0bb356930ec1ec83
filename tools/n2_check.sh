o=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err; wc -l $o/bench_n2.json; tail -c 700 $o/bench_n2.json
