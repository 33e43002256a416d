"""Bring-up / parity report for the tcgen05 implicit-GEMM convolutions.

Runs every (shape, op) case through libecgmm and through torch's fp32 convolution on the same
bf16-rounded operands, prints one line per case and exits non-zero if any case is outside
tolerance.  Used by tests/test_conv_gpu.py and directly under gpurun during bring-up.
"""
import os
import sys
import traceback
import zlib

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecgmm  # noqa: E402
from ecgmm import ops  # noqa: E402

DEV = "cuda"
# bf16 output rounding (2^-9 relative) on top of fp32 accumulation of bf16 products
RTOL, ATOL = 1.0 / 128, 2e-2


def ref_conv(x_nhwc, w_oihw, stride):
    """fp32 reference on bf16-rounded operands; returns NHWC fp32."""
    x = x_nhwc.float().permute(0, 3, 1, 2).contiguous()
    R, S = w_oihw.shape[2], w_oihw.shape[3]
    y = F.conv2d(x, w_oihw.to(torch.bfloat16).float(), None, stride, (R // 2, S // 2))
    return y.permute(0, 2, 3, 1).contiguous()


def err(a, b):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    tol = ATOL + RTOL * b.abs()
    return d.max().item(), (d / tol).max().item(), b.abs().max().item()


def run_case(name, N, H, W, Cin, Cout, R, S, stride, results, do=("fwd", "dgrad", "wgrad")):
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()) % (2**31))
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).to(DEV)
    w_fwd, w_dg = ops.conv_weight_prep(w)
    y_ref = ref_conv(x, w, stride)
    Ho, Wo = y_ref.shape[1], y_ref.shape[2]
    dy = torch.randn(N, Ho, Wo, Cout, generator=g).to(DEV).to(torch.bfloat16)
    for op in do:
        try:
            if op == "fwd":
                y = ops.conv2d_fwd(x, w_fwd, stride)
                torch.cuda.synchronize()
                e = err(y, y_ref)
            elif op == "dgrad":
                xf = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
                yy = F.conv2d(xf, w.to(torch.bfloat16).float(), None, stride, (R // 2, S // 2))
                yy.backward(dy.float().permute(0, 3, 1, 2).contiguous())
                dx_ref = xf.grad.permute(0, 2, 3, 1).contiguous()
                dx = ops.conv2d_dgrad(dy, w_dg, (H, W), stride)
                torch.cuda.synchronize()
                e = err(dx, dx_ref)
                # accumulate path: dx2 = dx0 + dgrad
                dx0 = torch.randn_like(dx_ref).to(torch.bfloat16)
                dx2 = dx0.clone()
                ops.conv2d_dgrad(dy, w_dg, (H, W), stride, out=dx2, accumulate=True)
                torch.cuda.synchronize()
                e2 = err(dx2, dx0.float() + dx_ref)
                e = (max(e[0], e2[0]), max(e[1], e2[1]), e[2])
            else:
                wf = w.to(torch.bfloat16).float().requires_grad_(True)
                xf = x.float().permute(0, 3, 1, 2).contiguous()
                yy = F.conv2d(xf, wf, None, stride, (R // 2, S // 2))
                yy.backward(dy.float().permute(0, 3, 1, 2).contiguous())
                dw_ref = wf.grad
                dw = torch.zeros_like(w)
                ops.conv2d_wgrad(x, dy, dw, R, S, stride)
                torch.cuda.synchronize()
                # fp32 accumulation of bf16 products: compare relative to the gradient scale
                d = (dw - dw_ref).abs().max().item()
                scale = dw_ref.abs().max().item()
                e = (d, d / (2e-3 * scale + 1e-6), scale)
            ok = e[1] <= 1.0
            results.append((name, op, ok, e))
            print(f"{'OK  ' if ok else 'FAIL'} {name:34s} {op:6s} maxabs={e[0]:.4g} worst/tol={e[1]:.3g} refmax={e[2]:.3g}",
                  flush=True)
        except Exception as ex:  # keep going: bring-up wants the full picture
            results.append((name, op, False, None))
            print(f"ERR  {name:34s} {op:6s} {type(ex).__name__}: {ex}", flush=True)
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as ex2:
                print("device unusable after error:", ex2, flush=True)
                return False
    return True


def run_stem(name, N, H, W, results):
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()) % (2**31))
    x = torch.randn(N, 3, H, W, generator=g).clamp(-1, 1).to(DEV)
    w = (torch.randn(64, 3, 7, 7, generator=g) / 147**0.5).to(DEV)
    xb = x.to(torch.bfloat16).float()
    wb = w.to(torch.bfloat16).float().requires_grad_(True)
    y_ref = F.conv2d(xb, wb, None, 2, 3)
    dy = torch.randn(y_ref.shape, generator=g).to(DEV).to(torch.bfloat16)
    y_ref.backward(dy.float())
    try:
        xs = ops.stem_s2d(x)
        ws = ops.stem_weight_prep(w)
        y = ops.stem_conv_fwd(xs, ws, H, W)
        torch.cuda.synchronize()
        e = err(y, y_ref.detach().permute(0, 2, 3, 1))
        ok = e[1] <= 1.0
        results.append((name, "fwd", ok, e))
        print(f"{'OK  ' if ok else 'FAIL'} {name:34s} fwd    maxabs={e[0]:.4g} worst/tol={e[1]:.3g} refmax={e[2]:.3g}",
              flush=True)
        dw = torch.zeros_like(w)
        ops.stem_conv_wgrad(xs, dy.permute(0, 2, 3, 1).contiguous(), dw, H, W)
        torch.cuda.synchronize()
        d = (dw - wb.grad).abs().max().item()
        scale = wb.grad.abs().max().item()
        e = (d, d / (2e-3 * scale + 1e-6), scale)
        ok = e[1] <= 1.0
        results.append((name, "wgrad", ok, e))
        print(f"{'OK  ' if ok else 'FAIL'} {name:34s} wgrad  maxabs={e[0]:.4g} worst/tol={e[1]:.3g} refmax={e[2]:.3g}",
              flush=True)
    except Exception as ex:
        results.append((name, "stem", False, None))
        print(f"ERR  {name:34s} stem   {type(ex).__name__}: {ex}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as ex2:
            print("device unusable after error:", ex2, flush=True)
            return False
    return True


CASES = [
    # name, N, H, W, Cin, Cout, R, S, stride
    ("3x3s1_64_64_small", 2, 10, 37, 64, 64, 3, 3, 1),
    ("1x1s1_64_64_gemm", 1, 8, 16, 64, 64, 1, 1, 1),
    ("3x3s1_128_128", 2, 9, 45, 128, 128, 3, 3, 1),
    ("3x3s2_64_128_odd", 2, 17, 45, 64, 128, 3, 3, 2),
    ("1x1s2_64_128_odd", 2, 17, 45, 64, 128, 1, 1, 2),
    ("3x3s2_128_256_even", 1, 16, 40, 128, 256, 3, 3, 2),
    ("3x3s1_256_256", 1, 16, 39, 256, 256, 3, 3, 1),
    ("3x3s2_256_512", 2, 16, 39, 256, 512, 3, 3, 2),
    ("3x3s1_512_512_8x79", 2, 8, 79, 512, 512, 3, 3, 1),
    ("1x3s1_64_64_1d", 3, 1, 619, 64, 64, 1, 3, 1),
    ("1x3s2_64_128_1d", 3, 1, 619, 64, 128, 1, 3, 2),
    ("1x1s2_128_256_1d", 2, 1, 310, 128, 256, 1, 1, 2),
    ("3x3s1_64_64_63x625", 2, 63, 625, 64, 64, 3, 3, 1),
    # weight gradient with N = 128 MMAs: two CTA types (taps 0..7 | tap 8 of a PAIR of Cin slices); odd slice counts
    ("3x3s1_64_128_oddslices", 2, 11, 50, 64, 128, 3, 3, 1),
    ("3x3s1_192_256_oddslices", 2, 9, 33, 192, 256, 3, 3, 1),
    ("3x3s1_128_128_32x313", 3, 32, 313, 128, 128, 3, 3, 1),
    ("1x3s1_128_128_1d", 3, 1, 310, 128, 128, 1, 3, 1),
    ("1x3s1_256_256_1d", 2, 1, 155, 256, 256, 1, 3, 1),
]


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ecgmm.lib.require_device()
    results = []
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    alive = True
    for c in CASES:
        if only and not any(o in c[0] for o in only):
            continue
        alive = run_case(*c, results=results)
        if not alive:
            break
    if alive and (not only or any("stem" in o for o in only)):
        for nm, N, H, W in (("stem_small", 2, 50, 100), ("stem_odd", 1, 37, 75), ("stem_250x2500", 2, 250, 2500)):
            if not run_stem(nm, N, H, W, results):
                break
    bad = [r for r in results if not r[2]]
    print(f"\n{len(results) - len(bad)}/{len(results)} cases OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
