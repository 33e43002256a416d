#!/bin/bash
set -u
TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
for f in test_attrib_serve_gpu test_conv_gpu test_data_gpu test_explain_gpu test_fusion_gpu; do
  run order_$f 300 python -m pytest tests/$f.py tests/test_graph_gpu.py -q -m gpu -k "not pair and not argmax"
done
run pytest 900 python -m pytest tests -q -m gpu --durations=5
run trace 300 python -m pytest tests/test_step_trace_gpu.py tests/test_dp_gpu.py -q -m gpu -s
cat $O/${TAG}_index.log
