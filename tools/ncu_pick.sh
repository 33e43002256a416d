#!/bin/bash
# `ncu --set full` of a HANDFUL of launches of the third training step at per-GPU batch 64 (tools/one_step.py), one
# short ncu run per kernel family; only the raw-page CSV of each run is kept (a full-step capture takes > 20 minutes
# and its report exceeds what gpurun copies back).   usage on the GPU box: bash tools/ncu_pick.sh r01c
set -u
TAG=${1:-r01c}
C="python tools/one_step.py 64 3"
$C > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
: > gpurun_out/${TAG}_full_raw.csv
pick() {  # name regex launches-to-skip launches-to-capture
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s "$3" -c "$4" -o gpurun_out/tmp_$1 $C \
    > gpurun_out/${TAG}_ncu_$1.log 2>&1
  if [ -f gpurun_out/tmp_$1.ncu-rep ]; then
    if [ -s gpurun_out/${TAG}_full_raw.csv ]; then
      ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 >> gpurun_out/${TAG}_full_raw.csv
    else
      ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv 2>/dev/null >> gpurun_out/${TAG}_full_raw.csv
    fi
    rm -f gpurun_out/tmp_$1.ncu-rep
  fi
}
# (--kernel-name-base demangled: the template arguments in the patterns below are part of the demangled name only;
#  without it the "<256", "<3, 0>" patterns of the first run matched nothing)
# per step: 17 wgrad_halo (first 3: layer4 with BN=128 types; last 4 of the image encoder: layer1), 12 halo, ...
# the whole weight-gradient class of the third step (bench.py's roofline object: 27 launches): DRAM traffic per launch
pick wgrad_all  "wgrad_halo_kernel|igemm_tn_kernel" 54 27
pick nt_halo    "igemm_nt_halo"       26 2
pick nt_256     "igemm_nt_kernel<256" 50 2
pick nt_128     "igemm_nt_kernel<128" 40 2
pick bn_bwd     "bn_bwd_apply_kernel<3, 0>|bn_bwd_reduce_kernel<3>" 100 2
pick stem       "stem_bwd_apply|stem_fwd_ring|bn_relu_maxpool" 6 3
pick bn_fwd     "bn_apply_kernel<0, 1, 1>|chan_stats" 66 2
rm -f gpurun_out/tmp_*.ncu-rep
ls -la gpurun_out/${TAG}_*
