"""Development probe (needs a B200): where do the product's stored activations first differ from the oracle evaluated
with bf16 rounding at the same tensors (tests/parity_util.storage_matched_oracle's forward)?  Prints, per stored tensor
of the image encoder, the fraction of elements that differ and the relative L2 difference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ecgmm  # noqa: E402,F401
from golden_util import make_inputs  # noqa: E402
from parity_util import build_pair  # noqa: E402

B, H, W, L = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (4, 64, 160, 600)))
ora, dut = build_pair(seed=7)
ora.train()
dut.train()
image, ecg, clin, labels = make_inputs(100 + B, B, H, W, L)
enc = ora.image_encoder
for mod in enc.modules():
    if isinstance(mod, torch.nn.ReLU):
        mod.inplace = False
cap = {}


def hook_for(name, rnd):
    def hook(mod, i, o):
        o2 = o.to(torch.bfloat16).float() if rnd else o
        cap[name] = o2.detach()
        return o2
    return hook


kinds = (torch.nn.Conv2d, torch.nn.BatchNorm2d, torch.nn.ReLU, torch.nn.MaxPool2d)
with torch.no_grad():
    for name, mod in enc.named_modules():
        if isinstance(mod, kinds):
            mod.register_forward_hook(hook_for(name, not name.endswith("bn2")))
        if isinstance(mod, torch.nn.Conv2d):
            mod.weight.copy_(mod.weight.to(torch.bfloat16).float())
    # block outputs (after the residual add + ReLU): BasicBlock.forward's final relu is the same module as the first one,
    # so capture at block level
    for lname in ("layer1", "layer2", "layer3", "layer4"):
        for bi, blk in enumerate(getattr(enc, lname)):
            blk.register_forward_hook(hook_for(f"{lname}.{bi}", True))
    enc(image.to(torch.bfloat16).float())

feat, state = dut.image_encoder.run_forward(image.cuda(), save=True)
xs, c1, st1, arg, recs, pooled, last_shape, hw = state


def cmp(tag, d, o_nchw):
    d = d.float().cpu()
    o = o_nchw.permute(0, 2, 3, 1).contiguous()
    diff = (d - o)
    frac = float((diff != 0).float().mean())
    rel = float(diff.norm() / o.norm())
    print(f"{tag:28s} differing elements {frac:9.6f}   rel L2 {rel:.3e}   max abs {float(diff.abs().max()):.3e}")


cmp("conv1 (c1)", c1, cap["conv1"])
cmp("stem pooled", recs[0][0], cap["maxpool"])
names = [f"layer{l}.{b}" for l in (1, 2, 3, 4) for b in (0, 1)]
for nm, rec in zip(names, recs):
    x, a, sa, m, b, sb, d, sd, mask_m, mask_out, se_rec = rec
    cmp(nm + ".conv1 (a)", a, cap[nm + ".conv1"])
    # m = relu(bn1(a)): the hooked relu module runs twice per block (after bn1 and at the end); cap holds the last call,
    # so compare m against relu(cap[bn1]) directly
    cmp(nm + ".relu(bn1) (m)", m, torch.relu(cap[nm + ".bn1"]))
    cmp(nm + ".conv2 (b)", b, cap[nm + ".conv2"])
    if d is not None:
        cmp(nm + ".downsample.0 (d)", d, cap[nm + ".downsample.0"])
for i, nm in enumerate(names[:-1]):
    cmp(nm + " output", recs[i + 1][0], cap[nm])
