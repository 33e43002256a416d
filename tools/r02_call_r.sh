#!/bin/bash
# 1 GPU: the one-kernel perturbation path (tests, A/B against the three-kernel path) and the nested-CV kfold bench
set -u
TAG=${1:-r02r}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 400 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-300))" | tee -a $O/${TAG}_index.log; }
run explain_tests 300 python -m pytest tests/test_explain_gpu.py tests/test_attrib_serve_gpu.py tests/test_modality_shapley_gpu.py -q -m gpu
run perturb_fused 300 python bench.py --config perturb
run perturb_3k 300 env ECGMM_PERTURB_FUSED=0 python bench.py --config perturb

cat $O/${TAG}_index.log
