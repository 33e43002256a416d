#!/bin/bash
set -u
TAG=${1:-r02p}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
run pytest 900 python -m pytest tests -q -m gpu -x
run b64    300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
run b512   300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
ECGMM_FUSED_BWD_REDUCE=0 run b64_nored  300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
ECGMM_FUSED_BWD_REDUCE=0 run b512_nored 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
cat $O/${TAG}_index.log
