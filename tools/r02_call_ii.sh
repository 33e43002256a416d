#!/bin/bash
# Last call of the round: cp.async-ring BatchNorm kernels (default on) - GPU suite, bench new vs ECGMM_BN_ASYNC=0 on the
# same box, smoke, per-kernel A/B.  Most important first: the call is cut off by the remaining GPU budget.
set -u
TAG=${1:-r02ii}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 120 python -m pytest tests -q -m gpu
run b512_async 60 python bench.py --no-cpu-baseline --steps 10
run b512_regs 60 env ECGMM_BN_ASYNC=0 python bench.py --no-cpu-baseline --steps 10
run smoke 60 python -c "import __graft_entry__ as g; g.smoke()"
run ab128 60 python tools/elementwise_ab.py --batch 128 --iters 10 --only-bn
