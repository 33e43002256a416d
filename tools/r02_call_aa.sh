#!/bin/bash
# batched weight prep: parity + the whole suite + bench at 64 and 512
set -u
TAG=${1:-r02aa}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 900 python -m pytest tests -q -m gpu -x
run bench_b64 200 python bench.py --global-batch 64 --no-cpu-baseline --steps 20
run bench_n1 300 python bench.py --no-cpu-baseline --steps 10
