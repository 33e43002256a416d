"""Runs the descriptor probe kernel (tools/csrc/desc_probe.cu) for every (layout, base-offset mode, row shift) and
reports which combinations reproduce the expected product.  Development tool (needs a B200); the probe is built into
its own shared library (tools/build/libecgmm_probe.so), it is not part of libecgmm.so or include/ecgmm.h.
Last result: profiles/r02_desc_probe.txt."""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import ecgmm  # noqa: E402
from ecgmm import lib as _lib, ops  # noqa: E402

_lib.require_device()
SO = os.path.join(HERE, "build", "libecgmm_probe.so")
if not os.path.exists(SO):
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    csrc = os.path.join(ROOT, "ecg-multimodal-model_b200", "csrc")
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-cudart", "static", "-shared",
                    "-I", os.path.join(ROOT, "include"), "-o", SO, os.path.join(HERE, "csrc", "desc_probe.cu"),
                    os.path.join(csrc, "common.cu")], check=True)
_so = ctypes.CDLL(SO)
_so.ecgmm_debug_desc_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p]


class lib:  # same call shape as ecgmm.lib.call
    @staticmethod
    def call(name, *args):
        rc = getattr(_so, name)(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed with status {rc}")


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
