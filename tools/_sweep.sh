for seg in 1 2 4 8 16; do
  echo "=== NH_SEG=$seg batch 64"
  ECGMM_NH_SEG=$seg python tools/conv_bench.py 64 8 layer1 2>&1 | grep -E "fwd|dgrad"
done
for seg in 1 4 16; do
  echo "=== NH_SEG=$seg batch 256"
  ECGMM_NH_SEG=$seg python tools/conv_bench.py 256 4 layer1 2>&1 | grep -E "fwd|dgrad"
done
