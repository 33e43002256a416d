#!/bin/bash
# 8 GPUs: the bench at N=8 (graph + NCCL), perturbation inference and nested-CV kfold sharded over 8 ranks, DP parity at 4
set -u
TAG=${1:-r02u}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 400 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-300))" | tee -a $O/${TAG}_index.log; }
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613"
run bench_n8 300 $T8 bench.py --gpus 8 --steps 10 --warmup 3
run perturb_n8 200 $T8 bench.py --config perturb
run kfold_n8 300 $T8 bench.py --config kfold

cat $O/${TAG}_index.log
