#!/bin/bash
# last build of the round: the whole GPU suite, smoke, bench at 512 and 64
set -u
TAG=${1:-r02final}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 900 python -m pytest tests -q -m gpu
run smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1 400 python bench.py
run bench_b64 200 python bench.py --global-batch 64 --no-cpu-baseline --steps 20
run perturb 200 python bench.py --config perturb
run signal 200 python bench.py --config signal
