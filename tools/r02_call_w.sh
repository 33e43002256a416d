#!/bin/bash
# CTA-pair variant of the one-kernel perturbation path: parity, then A/B against the single-CTA kernel
set -u
TAG=${1:-r02w}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run explain_tests 240 python -m pytest tests/test_explain_gpu.py -q -m gpu -x
run perturb_pair 200 python bench.py --config perturb
run perturb_single 200 env ECGMM_PERTURB_PAIR=0 python bench.py --config perturb
