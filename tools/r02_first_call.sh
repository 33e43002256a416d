#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent, in ONE gpurun
# call, each step under its own timeout so that a hanging experimental kernel cannot take the box with it.
#
#   gpurun --timeout 2700 -- 'bash tools/r02_first_call.sh r02a'      (about 30-40 GPU-minutes)
#
# Writes gpurun_out/<tag>_*.log.  Order: (1) the regular GPU suite (incl. the tests collected last that have never
# run on hardware), (2) smoke, (3) baseline bench lines at per-GPU batch 64 and 512, (4) the gated experiments, each
# first through its parity test and only then through the bench, (5) the per-config helpers.
set -u
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
run() {  # name timeout command...
  local name=$1 lim=$2
  shift 2
  echo "== $name: $*" | tee -a $O/${TAG}_index.log
  timeout "$lim" "$@" > $O/${TAG}_$name.log 2>&1
  echo "   rc=$? ($(tail -c 300 $O/${TAG}_$name.log | tr '\n' ' ' | cut -c1-200))" | tee -a $O/${TAG}_index.log
}
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"

run pytest        600 python -m pytest tests -x -q -m gpu
run smoke         300 python -c "import __graft_entry__ as g; g.smoke()"
run base_b64      300 $B --global-batch 64 --detail
run base_b512     300 $B --detail

# --- experiment 1: rolling-accumulator N=192 kernel for the 64-channel layers (csrc/conv_nt_stack.cu)
export ECGMM_TEST_EXPERIMENTAL=1
run stack_test    300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "rolling_accumulator"
if grep -q " passed" $O/${TAG}_stack_test.log && ! grep -q " failed" $O/${TAG}_stack_test.log; then
  ECGMM_NT_STACK=1 run stack_b64   300 $B --global-batch 64 --detail
  ECGMM_NT_STACK=1 run stack_b512  300 $B --detail
fi

# --- experiment 2: weight gradients on their own stream underneath the BatchNorm backward (model._WgradLane)
run wgstream_test 300 python -m pytest tests/test_fusion_gpu.py -x -q -m gpu -k "wgrad_lane"
if grep -q " passed" $O/${TAG}_wgstream_test.log && ! grep -q " failed" $O/${TAG}_wgstream_test.log; then
  ECGMM_WGRAD_STREAM=1 run wgstream_b64  300 $B --global-batch 64 --detail
  ECGMM_WGRAD_STREAM=1 run wgstream_b512 300 $B --detail
fi

# --- experiment 2b: transposed weight gradient (M = Cout, N = 192: wgrad_halo_kernel<128, true>)
run wgT_test 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "transposed_wgrad"
if grep -q " passed" $O/${TAG}_wgT_test.log && ! grep -q " failed" $O/${TAG}_wgT_test.log; then
  ECGMM_WG_T=1 run wgT_b64  300 $B --global-batch 64 --detail
  ECGMM_WG_T=1 run wgT_b512 300 $B --detail
fi

# --- experiment 3: folded-BatchNorm convolution epilogue of the serving path (ecgmm_conv2d_fwd_bn)
run fusedbn_test  300 python -m pytest tests/test_zz_attrib_serve_gpu.py -x -q -m gpu -k "experimental"
run serve_plain   300 python tools/serve_bench.py
if grep -q " passed" $O/${TAG}_fusedbn_test.log && ! grep -q " failed" $O/${TAG}_fusedbn_test.log; then
  ECGMM_SERVE_FUSED=1 run serve_fused 300 python tools/serve_bench.py
fi

# --- experiment 4: time-parallel signal preprocessing (signal_preprocess_block_kernel)
run prepblk_test  300 python -m pytest tests/test_preprocess_gpu.py -x -q -m gpu -k "block_parallel"
if grep -q " passed" $O/${TAG}_prepblk_test.log && ! grep -q " failed" $O/${TAG}_prepblk_test.log; then
  ECGMM_PREP_BLOCK=1 run signal_blk 300 python tools/signal_bench.py
fi
unset ECGMM_TEST_EXPERIMENTAL

# --- hardware question behind the transposed weight-gradient plan (DESIGN.md known headroom 3)
run desc_probe 120 python tools/desc_probe.py

# --- the other configs (BASELINE.json configs[1], [3], [4]) and the reference arm
run signal   300 python tools/signal_bench.py
run perturb  300 python tools/perturb_bench.py
ECGMM_TEST_EXPERIMENTAL=1 run zz_tests 600 python -m pytest tests/test_zz_modality_shapley_gpu.py tests/test_zz_attrib_serve_gpu.py -q -m gpu
run kfold    600 python tools/kfold_bench.py
run ref_arm  300 python bench.py --impl reference --steps 2 --warmup 1

# --- ncu launch list of the bench command itself (per-launch times are cold-cache and serialised: only each kernel's
# SHARE of the step is comparable with the CUDA-event numbers of the plain run above)
[ "${ECGMM_FIRSTCALL_NCU:-0}" = 1 ] && run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 6000 \
    --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --launch eager
cat $O/${TAG}_index.log
