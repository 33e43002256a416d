#!/bin/bash
# Round-2 profiles (run on the GPU box through gpurun; one ncu-using call):
#  (1) launch list of the bench command itself at per-GPU batch 64 (device time of every launch; cold-cache and
#      serialised, so only each kernel's SHARE of the step is comparable with the CUDA-event numbers);
#  (2) `ncu --set full` of a handful of launches of the third training step at batch 64 (tools/one_step.py), one short
#      run per kernel family; only the raw-page CSV of each run is kept (small), tools/ncu_summary.py turns it into the
#      tables under profiles/.                               usage: bash tools/ncu_pick_r02.sh r02
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --launch eager --global-batch 64"
$B > $O/${TAG}_ncu_plain_bench.log 2>&1 || { echo "plain bench run failed"; tail -5 $O/${TAG}_ncu_plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 8000 --csv \
  --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launches.log 2>&1
C="python tools/one_step.py 64 3"
$C > $O/${TAG}_ncu_plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 $O/${TAG}_ncu_plain_step.log; exit 1; }
: > $O/${TAG}_full_raw.csv
pick() {  # name regex launches-to-skip launches-to-capture
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s "$3" -c "$4" \
    -o $O/tmp_$1 $C > $O/${TAG}_ncu_$1.log 2>&1
  if [ -f $O/tmp_$1.ncu-rep ]; then
    if [ -s $O/${TAG}_full_raw.csv ]; then
      ncu -i $O/tmp_$1.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 >> $O/${TAG}_full_raw.csv
    else
      ncu -i $O/tmp_$1.ncu-rep --page raw --csv 2>/dev/null >> $O/${TAG}_full_raw.csv
    fi
    [ "$1" = "nt_pair" ] && cp $O/tmp_$1.ncu-rep $O/${TAG}_nt_pair.ncu-rep   # one report kept whole (source page)
    rm -f $O/tmp_$1.ncu-rep
  fi
}
# weight-gradient class of the third step (27 launches per step: skip the first two steps' launches)
pick wgrad_all  "wgrad_halo_kernel|igemm_tn_kernel" 54 27
pick nt_pair    "igemm_nt_pair_kernel" 52 4
pick nt_stack   "igemm_nt_stack_kernel" 16 4
pick stem       "stem_fwd_ring|stem_wgrad_ring|stem_bwd_apply|bn_relu_maxpool" 8 4
pick bn         "bn_apply_kernel|bn_bwd_apply_kernel|bn_bwd_reduce_kernel|chan_stats" 120 8
rm -f $O/tmp_*.ncu-rep
ls -la $O/${TAG}_*
