#!/bin/bash
# Final-build profiles (one ncu-using gpurun call, run only after the same commands exited 0 without ncu):
#  (1) launch list of the bench command at per-GPU batch 64 (cold-cache, serialised: compare SHARES);
#  (2) ONE light pass (5 metrics, no replay of the full set) over EVERY launch of the third training step of
#      tools/one_step.py at batch 64: time, DRAM bytes read / written, tensor-pipe %, DRAM % -> per-class traffic
#      (tools/ncu_traffic.py classifies the launches by kernel name and order);
#  (3) `ncu --set full` of two launches of the CTA-pair conv kernel and of the one-kernel perturbation path (the
#      kernels that are new since the mid-round full capture profiles/r02_ncu_full_summary.txt); raw-page CSV only.
# A first version captured `--set full` per class through NVTX ranges: 6 minutes per class (every pass restores the
# multi-GB activations the kernel writes) -- it ran into the call's time limit.
set -u
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --launch eager --global-batch 64"
$B > $O/${TAG}_ncu_plain_bench.log 2>&1 || { echo "plain bench run failed"; tail -5 $O/${TAG}_ncu_plain_bench.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 8000 --csv \
  --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launches.log 2>&1
echo "launch list: $(wc -l < $O/${TAG}_launches.csv) lines"
export ECGMM_SIDE_STREAM=0
C="python tools/one_step.py 64 3"
$C > $O/${TAG}_ncu_plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 $O/${TAG}_ncu_plain_step.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
ECGMM_PROFILE_LAST=1 timeout 400 ncu --profile-from-start off --metrics $M --clock-control none \
  --kernel-name-base demangled --csv --log-file $O/${TAG}_step_metrics.csv $C > $O/${TAG}_ncu_step.log 2>&1
echo "step metrics: $(wc -l < $O/${TAG}_step_metrics.csv) lines"
full() {  # name regex skip count command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$re" -s "$skip" -c "$cnt" \
    -o /tmp/ncu_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  [ -f /tmp/ncu_$name.ncu-rep ] && ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv 2>/dev/null > $O/${TAG}_full_$name.csv
  echo "$name: $(wc -l < $O/${TAG}_full_$name.csv 2>/dev/null) csv lines"
}
full nt_pair "igemm_nt_pair_kernel" 60 2 $C
unset ECGMM_SIDE_STREAM
P="python tools/perturb_bench.py --samples 64 --iters 2 --cpu-samples 1"
$P > $O/${TAG}_ncu_plain_perturb.log 2>&1 && full perturb_fused "perturb_fused_kernel" 3 1 $P
du -sh $O; ls -la $O/${TAG}_*
