#!/bin/bash
set -u
TAG=${1:-r02h}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
run kernels 600 python -m pytest tests/test_conv_gpu.py tests/test_kernels_gpu.py tests/test_step_trace_gpu.py tests/test_graph_gpu.py -q -m gpu -x -s
if grep -q " passed" $O/${TAG}_kernels.log && ! grep -q " failed" $O/${TAG}_kernels.log; then
  run b512 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
  ECGMM_FUSED_BWD_REDUCE=1 run b512_red 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
  run b64 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
  ECGMM_FUSED_BWD_REDUCE=1 run b64_red 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
fi
run pytest 900 python -m pytest tests -q -m gpu --durations=5
run signal 300 python bench.py --config signal
cat $O/${TAG}_index.log
