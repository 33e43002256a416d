#!/bin/bash
# One `ncu --set full` pass per kernel family over a short bench run; keeps only the raw CSV of each
# pass (the .ncu-rep files would exceed the 64 MiB that gpurun copies back).
# usage (on the GPU box, from the repo root): bash tools/ncu_pass.sh
set -u
C="python bench.py --global-batch 32 --steps 1 --no-cpu-baseline"
$C > gpurun_out/ncu_pass_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
pass() {  # name regex skip count
  ncu --set full --clock-control none -k regex:"$2" -s "$3" -c "$4" -o gpurun_out/tmp_$1 $C > gpurun_out/ncu_pass_$1.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/tmp_$1.ncu-rep
}
pass fwd_halo   "igemm_nt_halo"            0 2
pass fwd_nt     "igemm_nt_kernel"          0 12
pass wgrad      "wgrad_halo_kernel"        0 14
pass stem       "stem_fwd_ring|stem_wgrad_ring|stem_bwd|bn_relu_maxpool" 0 6
pass bn_fwd     "chan_stats|bn_apply"      0 8
pass bn_bwd     "bn_bwd_reduce|bn_bwd_apply" 40 18
ls -la gpurun_out/ncu_raw_*.csv
