#!/bin/bash
# 2-GPU validation: NCCL data-parallel parity (incl. oracle), the bench at N=2 (graph + NCCL), perturb / kfold sharded
set -u
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 400 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-300))" | tee -a $O/${TAG}_index.log; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
run dp_check 400 $T tools/dp_check.py
run bench_n2 400 $T bench.py --gpus 2 --steps 10 --warmup 3
run ref_n2 200 $T bench.py --impl reference --gpus 2 --steps 2 --warmup 1
run perturb_n2 300 $T bench.py --config perturb
run kfold_n2 400 $T bench.py --config kfold
cat $O/${TAG}_index.log
