"""Shared-memory operand-feed model of the convolution kernels against measured per-shape times (development helper).

    python tools/mma_model.py gpurun_out/s21_b64_detail.txt 64 > profiles/r01_mma_shape_model.txt

Model (known headroom 3): a tcgen05.mma M128 x N x K16 needs N/2 tensor-pipe clocks and reads (128 + N) * 32
bytes of shared-memory operands at 128 B/clk, so its rate is capped at  eff(N) = (N/2) / max(N/2, (128 + N)/4)  of the
pipe: 0.67 for N = 64, 1.0 (no slack) for N = 128, 1.0 for N = 256.  The reference rate for eff = 1 is the measured
sustained cuBLAS bf16 figure of MEASURED_PEAKS.json.  `quant` is the wave quantisation of the persistent grid (tiles /
(waves * 148 SMs)) for the forward / data-gradient tiles of 128 pixels x N channels.  measured / (model * quant) far
below 1 means the kernel loses time to something the model does not contain (pipeline bubbles, epilogue, split-K
reduction, L2); near 1 means only a different MMA shape can make it faster."""
import json
import math
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H0, W0 = 250, 2500


def eff(n):
    return (n / 2) / max(n / 2, (128 + n) / 4)


def feature_hw(cin):
    # input resolution of a ResNet18 stage at 250 x 2500 (stem /2, maxpool /2, then /2 per stage)
    h, w = (H0 - 1) // 2 + 1, (W0 - 1) // 2 + 1
    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    for c in (128, 256, 512):
        if cin >= c:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w


def main():
    path, batch = sys.argv[1], int(sys.argv[2])
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    print(f"# {path}, per-GPU batch {batch}; eff=1 rate {peak} TFLOP/s (cuBLAS bf16 sustained, MEASURED_PEAKS.json)")
    print(f"# {'kernel/shape':30s} {'n':>2s} {'ms':>7s} {'TFLOP/s':>8s} {'N':>4s} {'eff(N)':>6s} {'quant':>6s} {'model':>8s} {'meas/model':>10s}")
    pat = re.compile(r"# (conv_(fwd|dgrad|wgrad)/(\d+)x(\d+)k(\d)(\d)s(\d))\s+n=\s*(\d+)\s+([\d.]+) ms\s+([\d.]+) TFLOP/s")
    for line in open(path):
        m = pat.match(line)
        if not m:
            continue
        name, kind, cin, cout, r, s, st, n, ms, tf = m.groups()
        cin, cout, r, s, st, n, ms, tf = int(cin), int(cout), int(r), int(s), int(st), int(n), float(ms), float(tf)
        if r == 1 and s == 3:
            continue  # 1-D signal encoder: latency-bound, not a tensor-pipe question
        gemm_n = cout if kind == "fwd" else cin
        if kind == "wgrad":
            N = 64  # both weight-gradient kernels feed 64-wide Cout slices
            quant = 1.0
        else:
            N = 256 if gemm_n % 256 == 0 else (128 if gemm_n % 128 == 0 else 64)
            h, w = feature_hw(cin if kind == "fwd" else cout)  # forward: input grid; dgrad: dy grid
            oh, ow = (h, w) if kind == "dgrad" and st == 1 else (((h - 1) // st + 1, (w - 1) // st + 1) if kind == "fwd" else (h * st, w * st))
            if kind == "fwd":
                pix_h, pix_w = oh, ow
            else:
                # data gradient writes the INPUT grid of the layer = the grid one stage up for stride 2
                pix_h, pix_w = feature_hw(cout) if st == 1 else feature_hw(cin)
            tiles = batch * math.ceil(pix_h * pix_w / 128) * (gemm_n // N)
            waves = math.ceil(tiles / 148)
            quant = tiles / (waves * 148)
        model = peak * eff(N) * quant
        print(f"  {name:30s} {n:2d} {ms:7.3f} {tf:8.1f} {N:4d} {eff(N):6.2f} {quant:6.2f} {model:8.1f} {tf / model:10.2f}")


if __name__ == "__main__":
    main()
