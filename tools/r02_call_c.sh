#!/bin/bash
set -u
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
run bisect 300 python tools/forward_bisect.py
run pytest 900 python -m pytest tests -q -m gpu --durations=10
cat $O/${TAG}_index.log
