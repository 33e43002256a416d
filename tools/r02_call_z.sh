#!/bin/bash
# final build: conv parity (the large-split 1x1 weight-gradient fold changed), then the ncu passes
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_conv_gpu.py -q -m gpu -x > $O/r02z_conv_tests.log 2>&1
echo "conv tests rc=$? $(tail -1 $O/r02z_conv_tests.log)"
bash tools/ncu_traffic_r02.sh r02z
