#!/bin/bash
# 2 GPUs at 64 samples per GPU (the per-GPU batch of the 8-GPU run): what the gradient collectives cost next to the
# kernels, with NCCL's CTA count capped and with the buckets merged
set -u
TAG=${1:-r02v}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_$name.log | head -2 | tr '\n' ' '))" | tee -a $O/${TAG}_index.log; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
B="bench.py --gpus 2 --global-batch 128 --steps 20 --warmup 3 --no-cpu-baseline"
run one_gpu_b64 200 python bench.py --global-batch 64 --steps 20 --warmup 3 --no-cpu-baseline
run n2_default 200 $T $B
run n2_ctas4 200 env NCCL_MAX_CTAS=4 $T $B
run n2_ctas2 200 env NCCL_MAX_CTAS=2 $T $B
run n2_onebucket 200 env ECGMM_DP_MIN_BUCKET=1000000000 $T $B
run n2_onebucket_ctas4 200 env ECGMM_DP_MIN_BUCKET=1000000000 NCCL_MAX_CTAS=4 $T $B
cat $O/${TAG}_index.log
