"""Per-layer timing of the tcgen05 implicit-GEMM convolutions at the ResNet18 shapes of the
native 250x2500 configuration.  Development tool: prints one line
per (layer, op) with the CUDA-event time and the achieved TFLOP/s against MEASURED_PEAKS.json.

    python tools/conv_bench.py [batch] [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ecgmm  # noqa: E402,F401
from ecgmm import ops  # noqa: E402

# name, H, W (input), Cin, Cout, R, S, stride, multiplicity in ResNet18
LAYERS = [
    ("layer1.3x3", 63, 625, 64, 64, 3, 3, 1, 4),
    ("layer2.0.conv1", 63, 625, 64, 128, 3, 3, 2, 1),
    ("layer2.0.ds", 63, 625, 64, 128, 1, 1, 2, 1),
    ("layer2.3x3", 32, 313, 128, 128, 3, 3, 1, 3),
    ("layer3.0.conv1", 32, 313, 128, 256, 3, 3, 2, 1),
    ("layer3.0.ds", 32, 313, 128, 256, 1, 1, 2, 1),
    ("layer3.3x3", 16, 157, 256, 256, 3, 3, 1, 3),
    ("layer4.0.conv1", 16, 157, 256, 512, 3, 3, 2, 1),
    ("layer4.0.ds", 16, 157, 256, 512, 1, 1, 2, 1),
    ("layer4.3x3", 8, 79, 512, 512, 3, 3, 1, 3),
]


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # evict L2 between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    only_layer = sys.argv[3] if len(sys.argv) > 3 else None   # substring filter on the layer name
    only_op = sys.argv[4] if len(sys.argv) > 4 else None      # fwd | dgrad | wgrad
    ecgmm.lib.require_device()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    peak = peaks["bf16_tflops"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator().manual_seed(0)
    tot_ms, tot_flop = 0.0, 0.0
    print(f"batch {B}, peak {peak} TFLOP/s (burst, measured)")
    # stem
    H, W = 250, 2500
    x = torch.randn(B, 3, H, W, generator=g).clamp(-1, 1).cuda()
    xs = ops.stem_s2d(x)
    wst = ops.stem_weight_prep((torch.randn(64, 3, 7, 7, generator=g) / 12).cuda())
    Ho, Wo = 125, 1250
    dy = torch.randn(B, Ho, Wo, 64, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(64, 3, 7, 7, device="cuda")
    fl = 2.0 * B * Ho * Wo * 64 * 147
    for op, fn in (("fwd", lambda: ops.stem_conv_fwd(xs, wst, H, W)),
                   ("wgrad", lambda: ops.stem_conv_wgrad(xs, dy, dw, H, W))):
        if (only_layer and only_layer not in "stem") or (only_op and only_op != op):
            continue
        ms = time_fn(fn, iters, flush)
        print(f"{'stem 7x7':16s} {op:6s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s (algorithmic)  x1")
        tot_ms += ms
        tot_flop += fl
    del x, xs, dy
    for name, H, W, Cin, Cout, R, S, stride, mult in LAYERS:
        if only_layer and only_layer not in name:
            continue
        x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).cuda()
        w_fwd, w_dg = ops.conv_weight_prep(w)
        Ho, Wo = (H + 2 * (R // 2) - R) // stride + 1, (W + 2 * (S // 2) - S) // stride + 1
        dy = torch.randn(B, Ho, Wo, Cout, device="cuda").to(torch.bfloat16)
        dw = torch.zeros_like(w)
        dx = torch.empty_like(x)
        fl = 2.0 * B * Ho * Wo * Cout * Cin * R * S
        fns = {
            "fwd": lambda: ops.conv2d_fwd(x, w_fwd, stride),
            "dgrad": lambda: ops.conv2d_dgrad(dy, w_dg, (H, W), stride, out=dx),
            "wgrad": lambda: ops.conv2d_wgrad(x, dy, dw, R, S, stride),
        }
        for op, fn in fns.items():
            if only_op and only_op != op:
                continue
            ms = time_fn(fn, iters, flush)
            tf = fl / ms / 1e9
            print(f"{name:16s} {op:6s} {ms:8.3f} ms  {tf:8.1f} TFLOP/s  {100 * tf / peak:5.1f}% of measured peak  x{mult}",
                  flush=True)
            tot_ms += ms * mult
            tot_flop += fl * mult
        del x, dy, dx
    if tot_ms == 0:
        return
    print(f"TOTAL conv time per step (B={B}): {tot_ms:.2f} ms, {tot_flop / tot_ms / 1e9:.1f} TFLOP/s "
          f"-> conv-only ceiling {B / tot_ms * 1e3:.0f} samples/s")


if __name__ == "__main__":
    main()
