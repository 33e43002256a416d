"""ncu target: ONE launch of each stem / BatchNorm elementwise kernel at its largest shape of the training step
(per-GPU batch 64: stem 125x1250x64, layer1 63x625x64), after a warm-up launch outside the profiled region.

    ncu --profile-from-start off --set full --import-source on -k regex:"bn_|stem_bwd" -o out python tools/elementwise_once.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ecgmm  # noqa: E402,F401
from ecgmm import lib, ops  # noqa: E402

lib.require_device()
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
C = 64
x0 = torch.randn(N, 125, 1250, C, device=dev).to(torch.bfloat16)          # conv1 output
x1 = torch.randn(N, 63, 625, C, device=dev).to(torch.bfloat16)            # a layer1 conv output
r1 = torch.randn_like(x1)
d1 = torch.randn_like(x1)
gamma, beta = torch.randn(C, device=dev) * 0.5 + 0.2, torch.randn(C, device=dev) * 0.1
st0 = ops.bn_train_stats(x0, gamma, beta, None, None, None, 1e-5, 0.1)
st1 = ops.bn_train_stats(x1, gamma, beta, None, None, None, 1e-5, 0.1)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)


def work():
    y, arg = ops.bn_relu_maxpool(x0, st0)
    dy = torch.ones_like(y)
    ops.bn_backward(x0, dy, st0, gamma, argmax=arg, pooled=y, beta=beta, dgamma=dg, dbeta=db)   # m4 reduce + rows apply
    _, m = ops.bn_apply(x1, st1, relu=True, want_mask=True)
    ops.bn_apply(x1, st1, res=r1, relu=True, want_mask=True)
    ops.bn_backward(x1, d1, st1, gamma, mask=m, want_dz=True, dgamma=dg, dbeta=db)               # mode-3 reduce + apply(+dz)
    ops.bn_backward(x1, d1, st1, gamma, dgamma=dg, dbeta=db)                                     # mode-0 reduce + apply


work()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
work()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
