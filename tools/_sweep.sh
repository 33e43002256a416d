set -x
python -m pytest tests/test_conv_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
for cfg in "128 1" "128 2" "64 3" "64 2" "96 2" "48 4" "32 7" "112 2" "80 2"; do
  set -- $cfg
  echo "=== KP=$1 RPS=$2"
  ECGMM_WG_KP=$1 ECGMM_WG_RPS=$2 python tools/conv_bench.py 64 5 3x3 wgrad 2>&1 | grep wgrad
done
echo "=== halo fwd/dgrad rolling"
python tools/conv_bench.py 64 5 layer1 2>&1 | grep layer1
