#!/bin/bash
set -u
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
run() { local name=$1 lim=$2; shift 2; echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-220))" | tee -a $O/${TAG}_index.log; }
ECGMM_NT_PAIR=1 run pair_check 180 python tests/gpu_conv_check.py 128_128 256_256 512_512 128_256 256_512 64_128
run pair_tests 400 python -m pytest tests/test_conv_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "pair or epilogue_statistics" -x
if grep -q " passed" $O/${TAG}_pair_tests.log && ! grep -q " failed" $O/${TAG}_pair_tests.log; then
  ECGMM_NT_PAIR=1 run b512_pair 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
  run b512_base 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --detail
  ECGMM_NT_PAIR=1 run b64_pair 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
  run b64_base 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 --detail
fi
CUDA_LAUNCH_BLOCKING=1 run pytest_blocking 900 python -m pytest tests -q -m gpu -x --tb=long -k "not pair"
run signal 300 python bench.py --config signal
cat $O/${TAG}_index.log
