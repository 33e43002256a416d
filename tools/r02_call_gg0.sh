#!/bin/bash
# ncu --set full of one launch of each stem / BatchNorm elementwise kernel (tools/elementwise_once.py), after the same
# command exited 0 without ncu.
set -u
TAG=${1:-r02gg0}
O=gpurun_out
mkdir -p $O
E="python tools/elementwise_once.py 64"
timeout 100 $E > $O/${TAG}_ncu_plain_once.log 2>&1 || { echo "plain elementwise_once run failed"; tail -5 $O/${TAG}_ncu_plain_once.log; exit 1; }
timeout 240 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k regex:"bn_relu_maxpool|stem_bwd_apply|bn_apply|bn_bwd_apply|bn_bwd_reduce" -c 10 -o $O/${TAG}_elementwise $E \
  > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log
ls -la $O/${TAG}_*
