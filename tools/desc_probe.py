"""Runs ecgmm_debug_desc_probe for every (layout, base-offset mode, row shift) and reports which
combinations reproduce the expected product.  Development tool (needs a B200)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecgmm  # noqa: E402
from ecgmm import lib, ops  # noqa: E402

lib.require_device()
g = torch.Generator().manual_seed(0)
a = torch.randn(160, 128, generator=g).cuda().to(torch.bfloat16)
for mn in (0, 1):
    b = torch.randn(32 if mn else 64, 64, generator=g).cuda().to(torch.bfloat16)
    for boff in (0, 1):
        oks = []
        for shift in range(0, 17):
            out = torch.full((128, 64), float("nan"), device="cuda")
            lib.call("ecgmm_debug_desc_probe", ops._ptr(a), ops._ptr(b), ops._ptr(out), shift, mn | (boff << 1),
                     ops._s())
            torch.cuda.synchronize()
            if mn:
                ref = a[shift:shift + 32].float().t() @ b.float()           # [128 m][64 n]
            else:
                ref = a[shift:shift + 128, :64].float() @ b.float().t()     # [128 m][64 n]
            err = (out - ref).abs().max().item()
            oks.append(err < 1e-2 * ref.abs().max().item())
        print(f"layout={'MN' if mn else 'K '}-major base_offset={'set' if boff else '0  '}:",
              "".join("Y" if o else "." for o in oks), "(shift 0..16)")

# mode bit 2: overlapping N atoms (leading-dimension offset = one 128-byte row) in an MN-major B operand, N = 192
for boff in (0, 1):
    oks = []
    for shift in range(0, 17):
        out = torch.full((128, 192), float("nan"), device="cuda")
        lib.call("ecgmm_debug_desc_probe", ops._ptr(a), ops._ptr(a), ops._ptr(out), shift, 4 | 1 | (boff << 1),
                 ops._s())
        torch.cuda.synchronize()
        af = a.float()
        ref = torch.cat([af[0:32].t() @ af[shift + j:shift + j + 32, :64] for j in range(3)], dim=1)  # [128][192]
        err = (out - ref).abs().max().item()
        oks.append(err < 1e-2 * ref.abs().max().item())
    print(f"MN-major B, N=192, atoms 128 B apart, base_offset={'set' if boff else '0  '}:",
          "".join("Y" if o else "." for o in oks), "(shift 0..16)")
