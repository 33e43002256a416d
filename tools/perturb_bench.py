"""configs[3] on one B200: 4096 masked variants per sample through the fusion head (development / profiles helper).

    python tools/perturb_bench.py [--samples 256] [--variants 4096] [--cpu-samples 2]

Prints one JSON line: samples/s and variants/s of ecgmm.explain.perturbation_inference (device-resident inputs,
CUDA events), the per-kernel split (variant build: HBM-bound, GEMM: tensor-bound, tail: HBM-bound) with achieved
GB/s / TFLOP/s, and the fp32 oracle on the host cores on a bounded number of samples."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--variants", type=int, default=4096)
    ap.add_argument("--cpu-samples", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    import torch

    import ecgmm
    from ecgmm import explain, lib
    from oracle import model as om

    lib.require_device()
    dev = torch.device("cuda", 0)
    S, V, D, HID = args.samples, args.variants, 768, 128

    class Cfg:
        num_classes = 2
        device = dev

    torch.manual_seed(42)
    model = ecgmm.ECGMultimodalModel(Cfg).eval()
    head = model.fusion_classifier
    g = torch.Generator().manual_seed(42)
    e = torch.randn(S, D, generator=g)
    bg = torch.randn(100, D, generator=g).mean(0)
    masks = (torch.rand(V, D, generator=g) < 0.5).to(torch.uint8)
    ed, bd, md = e.to(dev), bg.to(dev), masks.to(dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for _ in range(3):
        explain.perturbation_inference(head, ed, bd, md, 1)
    torch.cuda.synchronize()
    t = {"build": 0.0, "gemm": 0.0, "tail": 0.0}
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.iters):
        out = explain.perturbation_inference(head, ed, bd, md, 1)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / args.iters
    # per-kernel split (same calls, bracketed individually)
    from ecgmm import ops

    w1, b1, w2, b2 = explain._head_weights(head)
    for _ in range(args.iters):
        a, b, c, d = ev(), ev(), ev(), ev()
        a.record()
        x = explain.masked_variants(ed, bd, md)
        b.record()
        hidden = ops.conv2d_fwd(x.view(1, 1, S * V, D), w1, 1)
        c.record()
        o = torch.empty(S * V, dtype=torch.float32, device=dev)
        lib.call("ecgmm_head_tail", ops._ptr(hidden), ops._ptr(b1), ops._ptr(w2), ops._ptr(b2), ops._ptr(o), S * V,
                 HID, 2, 1, ops._s())
        d.record()
        torch.cuda.synchronize()
        t["build"] += a.elapsed_time(b) / args.iters
        t["gemm"] += b.elapsed_time(c) / args.iters
        t["tail"] += c.elapsed_time(d) / args.iters
    rows = S * V
    flops = 2.0 * rows * D * HID
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks = json.load(open(pp))
    kern = {
        "perturb_build": {"ms": round(t["build"], 4), "bound": "hbm",
                          "achieved_GBs": round((rows * D * 2 + S * D * 4) / (t["build"] * 1e-3) / 1e9, 1)},
        "gemm_1x1_tcgen05": {"ms": round(t["gemm"], 4), "bound": "tensor",
                             "achieved_TFLOPs": round(flops / (t["gemm"] * 1e-3) / 1e12, 1)},
        "head_tail": {"ms": round(t["tail"], 4), "bound": "hbm",
                      "achieved_GBs": round((rows * HID * 2 + rows * 4) / (t["tail"] * 1e-3) / 1e9, 1)},
    }
    kern["perturb_build"]["frac"] = round(kern["perturb_build"]["achieved_GBs"] / peaks["hbm_gbs"], 3)
    kern["gemm_1x1_tcgen05"]["frac"] = round(kern["gemm_1x1_tcgen05"]["achieved_TFLOPs"] / peaks["bf16_tflops_sustained"], 3)
    kern["head_tail"]["frac"] = round(kern["head_tail"]["achieved_GBs"] / peaks["hbm_gbs"], 3)
    # CPU oracle on a bounded sample
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ora = om.ECGMultimodalModel().eval()
    ns = max(1, args.cpu_samples)
    om.perturbation_inference(ora.fusion_classifier, e[:1], bg, masks, 1)
    t0 = time.perf_counter()
    ref = om.perturbation_inference(ora.fusion_classifier, e[:ns], bg, masks, 1)
    cpu_s = time.perf_counter() - t0
    line = {"metric": "perturbation-inference samples/sec (4096 variants each)", "value": S / (total_ms * 1e-3),
            "unit": "samples/s", "variants_per_s": rows / (total_ms * 1e-3), "ms_per_call": total_ms, "n_gpus": 1,
            "config": {"workload": "configs[3]: masked variants through fusion_classifier", "samples": S, "variants": V,
                       "D": D, "hidden": HID}, "dtype": "bf16", "kernels": kern,
            "cpu_baseline": {"value": ns / cpu_s, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{ns} samples x {V} variants through oracle.model.perturbation_inference (fp32)"}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
