#!/bin/bash
# timing-only ablations of the CTA-pair perturbation kernel (results are wrong by construction)
O=gpurun_out; mkdir -p $O
for f in 0 1 2 4 8 16 3 7 15; do
  ECGMM_PF_DEBUG=$f timeout 120 python tools/perturb_bench.py --cpu-samples 1 --iters 10 > $O/r02y_$f.log 2>&1
  echo "flags=$f rc=$? $(grep -o '"perturb_fused": {"ms": [0-9.]*' $O/r02y_$f.log)"
done
