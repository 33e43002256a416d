#!/bin/bash
# tiled stem max-pool: parity + bench A/B
set -u
TAG=${1:-r02bb}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pool_tests 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "pool or stem"
run bench_new 300 python bench.py --no-cpu-baseline --steps 10
run bench_old 300 env ECGMM_POOL_LEGACY=1 python bench.py --no-cpu-baseline --steps 10
