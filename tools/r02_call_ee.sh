#!/bin/bash
# A/B of the last elementwise changes of the round: branch-free tiled max-pool, row-wise stem_bwd_apply, clustered
# BatchNorm finalize folds.  GPU suite on the new defaults, per-kernel A/B, bench at batch 64 and 512 new vs old.
set -u
TAG=${1:-r02ee}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
OLD="env ECGMM_POOL_TILED_V1=1 ECGMM_STEM_BWD_APPLY=0 ECGMM_FIN_CLUSTER=0"
run pytest_gpu 600 python -m pytest tests -q -m gpu -x
run ab64 120 python tools/elementwise_ab.py --batch 64
run ab256 120 python tools/elementwise_ab.py --batch 256 --iters 10
run b64_new 150 python bench.py --global-batch 64 --no-cpu-baseline --steps 30
run b64_old 150 $OLD python bench.py --global-batch 64 --no-cpu-baseline --steps 30
run b512_new 200 python bench.py --no-cpu-baseline
run b512_old 200 $OLD python bench.py --no-cpu-baseline
run b64_new2 150 python bench.py --global-batch 64 --no-cpu-baseline --steps 30
