#!/bin/bash
# Round-1 profile of ONE training step at per-GPU batch 64 (the 8-GPU configuration's shard):
#   1. launch list (gpu__time_duration.sum, --clock-control none) of the whole bench run -> rNN_launches.csv
#   2. `ncu --set full` of every convolution / BatchNorm / stem launch of one steady-state step -> raw CSV
# usage (on the GPU box, from the repo root): bash tools/ncu_step.sh r01c
set -u
TAG=${1:-r01c}
C="python bench.py --global-batch 64 --steps 1 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/${TAG}_launches.csv $C \
  > gpurun_out/${TAG}_ncu_launches.log 2>&1
K="wgrad_halo_kernel|igemm_nt|igemm_tn|bn_bwd_reduce|bn_bwd_apply|stem_fwd|stem_wgrad|stem_bwd|bn_relu_maxpool|chan_stats|bn_apply"
ncu --set full --clock-control none -k regex:"$K" -s ${2:-630} -c ${3:-210} -o gpurun_out/tmp_${TAG} $C \
  > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/tmp_${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
rm -f gpurun_out/tmp_${TAG}.ncu-rep
ls -la gpurun_out/${TAG}_*
