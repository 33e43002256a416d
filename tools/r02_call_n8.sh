#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $T8 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02final_bench_n8.log 2>&1
echo "rc=$? $(grep -o '"value": [0-9.]*' $O/r02final_bench_n8.log | head -2 | tr '\n' ' ')"
