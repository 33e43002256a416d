#!/bin/bash
# Final-build profiles (one ncu-using gpurun call, run only after the same commands exited 0 without ncu):
#  (1) launch list of the bench command at per-GPU batch 64 (cold-cache, serialised: compare SHARES);
#  (2) `ncu --set full` of ONE training step's launches of each dominant roofline class (third step of tools/one_step.py
#      at batch 64, selected through the NVTX ranges ecgmm.ops opens with ECGMM_NVTX=1), for roofline.traffic;
#  (3) `ncu --set full` of the one-kernel perturbation path (configs[3]).
# Only the raw-page CSVs are kept; tools/ncu_traffic.py turns them into profiles/r02_traffic.json + a table.
set -u
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --launch eager --global-batch 64"
$B > $O/${TAG}_ncu_plain_bench.log 2>&1 || { echo "plain bench run failed"; tail -5 $O/${TAG}_ncu_plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 8000 --csv \
  --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launches.log 2>&1
export ECGMM_NVTX=1 ECGMM_SIDE_STREAM=0
C="python tools/one_step.py 64 3"
$C > $O/${TAG}_ncu_plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 $O/${TAG}_ncu_plain_step.log; exit 1; }
pick() {  # class launches-per-step
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled --nvtx --nvtx-include "$1]" \
    -s $((2 * $2)) -c "$2" -o $O/tmp_$1 $C > $O/${TAG}_ncu_$1.log 2>&1
  if [ -f $O/tmp_$1.ncu-rep ]; then
    ncu -i $O/tmp_$1.ncu-rep --page raw --csv 2>/dev/null > $O/${TAG}_full_$1.csv
    [ "$1" = "conv_fwd" ] && cp $O/tmp_$1.ncu-rep $O/${TAG}_conv_fwd.ncu-rep   # one report kept whole (source page)
    rm -f $O/tmp_$1.ncu-rep
  fi
  echo "$1: $(wc -l < $O/${TAG}_full_$1.csv 2>/dev/null) csv lines; $(tail -2 $O/${TAG}_ncu_$1.log | tr '\n' ' ')"
}
# launches per training step of the fusion model at 250x2500 (kernels.<class>.launches of the bench line; a weight
# gradient with a split-K workspace is two kernels per call, the stem-side classes are separate)
pick conv_fwd 28
pick conv_dgrad 27
pick conv_wgrad 56
pick bn_bwd_apply 29
pick bn_bwd_reduce 29
pick bn_apply 27
unset ECGMM_NVTX ECGMM_SIDE_STREAM
P="python tools/perturb_bench.py --samples 64 --iters 2"
$P > $O/${TAG}_ncu_plain_perturb.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:perturb_fused_kernel -s 3 -c 1 -o $O/tmp_pf $P > $O/${TAG}_ncu_perturb.log 2>&1
if [ -f $O/tmp_pf.ncu-rep ]; then
  ncu -i $O/tmp_pf.ncu-rep --page raw --csv 2>/dev/null > $O/${TAG}_full_perturb_fused.csv
  cp $O/tmp_pf.ncu-rep $O/${TAG}_perturb_fused.ncu-rep
fi
rm -f $O/tmp_*.ncu-rep
ls -la $O/${TAG}_*
