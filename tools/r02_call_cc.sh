#!/bin/bash
# transposed weight-gradient fold with 4x the CTAs: parity, then the threshold between the tiled and the per-element fold
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x > $O/r02cc_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/r02cc_tests.log)"
for mk in 6 12 24 64 0; do
  ECGMM_WG_REDUCE_T_MAXKS=$mk timeout 200 python bench.py --global-batch 64 --no-cpu-baseline --steps 30 > $O/r02cc_b64_$mk.log 2>&1
  echo "maxks=$mk $(grep -o '"ms_per_step": [0-9.]*' $O/r02cc_b64_$mk.log | head -1)"
done
