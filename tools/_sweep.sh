echo "=== halo fwd N=64 (normal)"
python tools/conv_bench.py 64 8 layer1 fwd 2>&1 | grep -E "fwd"
echo "=== halo fwd EXPERIMENT N=128 MMAs (2x the MMA flops, half of them garbage)"
ECGMM_EXP_N128=1 python tools/conv_bench.py 64 8 layer1 fwd 2>&1 | grep -E "fwd"
