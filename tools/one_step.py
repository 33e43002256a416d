"""Three training steps of the fusion model at a given per-GPU batch and nothing else (profiling target: ncu
attaches to this instead of the whole bench.py).   python tools/one_step.py [batch=64] [steps=3]
ECGMM_PROFILE_LAST=1: cudaProfilerStart() before the last step (for `ncu --profile-from-start off`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ecgmm  # noqa: E402
from ecgmm import nn as enn  # noqa: E402
from ecgmm import optim as eoptim  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)


class Cfg:
    num_classes = 2
    device = dev


torch.manual_seed(42)
model = ecgmm.ECGMultimodalModel(Cfg).train()
crit = enn.CrossEntropyLoss()
opt = eoptim.Adam(model.parameters(), lr=1e-4)
image, ecg, clin, labels = [t.to(dev) for t in bench.synth_batch(B, 42)]
for i in range(steps):
    if i == steps - 1 and os.environ.get("ECGMM_PROFILE_LAST") == "1":  # ncu --profile-from-start off: last step only
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    opt.zero_grad()
    out = model(image, ecg, clin)
    (crit(out[3], labels) + 0.1 * out[4]).backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(out[3].float().abs().sum()))
