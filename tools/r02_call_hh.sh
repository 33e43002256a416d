#!/bin/bash
# Shipping build of the round (call gg's build minus the two slower BatchNorm variants): GPU suite, smoke, the driver's bench command.
set -u
TAG=${1:-r02hh}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 300 python -m pytest tests -q -m gpu
run smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1 200 python bench.py --gpus 1 --steps 20 --warmup 5
