#!/bin/bash
# Final build of the round: GPU suite, smoke, bench lines (driver's K/W), then the ncu evidence of the SAME build: a
# light per-launch pass (time, DRAM bytes, tensor pipe %, DRAM %) over every launch of one training step at per-GPU
# batch 64 = launch list + per-class DRAM traffic.  (`--set full` of the changed elementwise kernels: r02_call_gg0.sh.)
set -u
TAG=${1:-r02gg}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-250))" | tee -a $O/${TAG}_index.log; }
run pytest_gpu 400 python -m pytest tests -q -m gpu
run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1 300 python bench.py --gpus 1 --steps 20 --warmup 5
run bench_b64 120 python bench.py --global-batch 64 --no-cpu-baseline --steps 30 --detail
run signal 120 python bench.py --config signal
run ab128 150 python tools/elementwise_ab.py --batch 128 --iters 10
run b512_fast2 150 env ECGMM_BN_FAST=2 python bench.py --no-cpu-baseline --steps 10
run b512_fast3 150 env ECGMM_BN_FAST=3 python bench.py --no-cpu-baseline --steps 10
run b512_fast1 150 python bench.py --no-cpu-baseline --steps 10
# ---- ncu (numbers printed under ncu are never bench values)
export ECGMM_SIDE_STREAM=0
C="python tools/one_step.py 64 3"
timeout 100 $C > $O/${TAG}_ncu_plain_step.log 2>&1 || { echo "plain step run failed"; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
ECGMM_PROFILE_LAST=1 timeout 200 ncu --profile-from-start off --metrics $M --clock-control none \
  --kernel-name-base demangled --csv --log-file $O/${TAG}_step_metrics.csv $C > $O/${TAG}_ncu_step.log 2>&1
echo "step metrics: $(wc -l < $O/${TAG}_step_metrics.csv) lines" | tee -a $O/${TAG}_index.log
ls -la $O/${TAG}_* | awk '{print $5, $9}'
