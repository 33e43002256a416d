#!/bin/bash
# 40 seconds of GPU budget left: the shipping library (call ii's build minus the forward-apply cp.async kernel) through smoke()
set -u
O=gpurun_out
mkdir -p $O
timeout 35 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02jj_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $O/r02jj_smoke.log | cut -c1-200)"
