#!/bin/bash
# generic (stride-2 / 1x1) weight-gradient fold with 4x the CTAs: parity, then the threshold sweep at batch 64
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x > $O/r02dd_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r02dd_tests.log)"
for mk in 6 16 40 200 0; do
  ECGMM_TN_REDUCE_MAXKS=$mk timeout 200 python bench.py --global-batch 64 --no-cpu-baseline --steps 30 > $O/r02dd_b64_$mk.log 2>&1
  echo "maxks=$mk $(grep -o '"ms_per_step": [0-9.]*' $O/r02dd_b64_$mk.log | head -1)"
done
