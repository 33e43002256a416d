#!/bin/bash
# timing-only ablations of the CTA-pair perturbation kernel (results are wrong by construction).  Kept for the record:
# the ECGMM_PF_DEBUG switches (skip fence / st.shared / e,b loads / epilogue arithmetic / MMA) lived in a scratch build of
# csrc/perturb_fused.cu and were removed again; the numbers are in profiles/r02y_perturb_ablation.txt.
O=gpurun_out; mkdir -p $O
for f in 0 1 2 4 8 16 3 7 15; do
  ECGMM_PF_DEBUG=$f timeout 120 python tools/perturb_bench.py --cpu-samples 1 --iters 10 > $O/r02y_$f.log 2>&1
  echo "flags=$f rc=$? $(grep -o '"perturb_fused": {"ms": [0-9.]*' $O/r02y_$f.log)"
done
