#!/bin/bash
# A/B of the BatchNorm fast paths (register-resident coefficients, packed ReLU mask) on top of call ee's winners.
set -u
TAG=${1:-r02ff}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 600 python -m pytest tests -q -m gpu -x
run ab64 150 python tools/elementwise_ab.py --batch 64
run ab256 150 python tools/elementwise_ab.py --batch 256 --iters 10
run b512_new 200 python bench.py --no-cpu-baseline
run b512_old 200 env ECGMM_BN_FAST=0 python bench.py --no-cpu-baseline
run b64_new 150 python bench.py --global-batch 64 --no-cpu-baseline --steps 30
run b64_old 150 env ECGMM_BN_FAST=0 python bench.py --global-batch 64 --no-cpu-baseline --steps 30
