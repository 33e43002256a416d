#!/bin/bash
# source-level stall attribution of the one-kernel perturbation path (single-CTA and CTA-pair variants)
set -u
O=gpurun_out
mkdir -p $O
P="python tools/perturb_bench.py --samples 128 --iters 2 --cpu-samples 1"
for mode in 0 1; do
  export ECGMM_PERTURB_PAIR=$mode
  $P > $O/r02x_plain_$mode.log 2>&1 || { echo "plain run failed ($mode)"; tail -3 $O/r02x_plain_$mode.log; continue; }
  timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on \
    --clock-control none --kernel-name-base demangled -k regex:perturb_fused -s 3 -c 1 -o /tmp/pf_$mode $P > $O/r02x_ncu_$mode.log 2>&1
  if [ -f /tmp/pf_$mode.ncu-rep ]; then
    ncu -i /tmp/pf_$mode.ncu-rep --page source --csv 2>/dev/null > $O/r02x_source_$mode.csv
    ncu -i /tmp/pf_$mode.ncu-rep --page raw --csv 2>/dev/null > $O/r02x_raw_$mode.csv
  fi
  echo "mode $mode: $(wc -l < $O/r02x_source_$mode.csv 2>/dev/null) source lines"
done
du -sh $O
