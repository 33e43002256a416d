python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5
echo "=== BN=128 forced (ECGMM_WG_BN=128)"
ECGMM_WG_BN=128 python tools/conv_bench.py 64 5 3x3 wgrad 2>&1 | grep wgrad
echo "=== BN=64 forced (ECGMM_WG_BN=64)"
ECGMM_WG_BN=64 python tools/conv_bench.py 64 5 3x3 wgrad 2>&1 | grep wgrad
