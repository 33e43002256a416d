#!/bin/bash
set -u
TAG=${1:-r02i}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
run pytest 900 python -m pytest tests -q -m gpu --durations=5
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run perturb 300 python bench.py --config perturb
run kfold 600 python bench.py --config kfold
run ncu 2400 bash tools/ncu_pick_r02.sh r02
cat $O/${TAG}_index.log
