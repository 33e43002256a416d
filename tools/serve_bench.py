"""Serving path on one B200 (development / profiles helper).

    python tools/serve_bench.py [--batch 64] [--iters 20] [--cpu-images 4]
    ECGMM_SERVE_FUSED=1 python tools/serve_bench.py          # folded BatchNorm applied by the convolution epilogues

Prints one JSON line: images/s of ecgmm.serve.ImageEndpoint at 3x250x2500 uint8 input (device-resident requests,
CUDA events) for the eager kernel sequence, the one-launch CUDA graph and the graph with Grad-CAM; `e2e` = pinned host
uint8 batch copied H2D every request + probabilities read back; and the fp32 oracle (oracle.model.image_endpoint) on the
host cores on a bounded number of images."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H, W = 250, 2500


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-images", type=int, default=4)
    args = ap.parse_args()
    import torch

    import ecgmm
    from ecgmm import lib, serve
    from oracle import model as om

    lib.require_device()
    dev = torch.device("cuda", 0)
    B = args.batch

    class Cfg:
        num_classes = 2
        device = dev

    torch.manual_seed(42)
    model = ecgmm.ECGMultimodalModel(Cfg).eval()
    g = torch.Generator().manual_seed(42)
    host = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).pin_memory()
    req = host.to(dev)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    eager = serve.ImageEndpoint(model, graph=False)
    graphed = serve.ImageEndpoint(model, example_image=req, graph=True)
    n0 = lib.launch_count()
    eager(req)
    launches = lib.launch_count() - n0
    ms_eager = timed(lambda: eager(req), args.iters)
    ms_graph = timed(lambda: graphed(graphed.input), args.iters)
    ms_cam = timed(lambda: graphed.gradcam(graphed.input), args.iters)

    def e2e():
        probs, _ = graphed(host.to(dev, non_blocking=True))
        return probs.cpu()

    ms_e2e = timed(e2e, args.iters)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ora = om.ECGMultimodalModel().eval()
    n = max(1, args.cpu_images)
    img = (host[:n].float() / 255.0 - 0.5) / 0.5
    om.image_endpoint(ora, img[:1])
    t0 = time.perf_counter()
    om.image_endpoint(ora, img)
    cpu_s = time.perf_counter() - t0
    fwd_gflop = 46.194  # ResNet18 forward at 250x2500 per image
    line = {"metric": "image-endpoint images/sec", "value": B / (ms_graph * 1e-3), "unit": "images/s", "n_gpus": 1,
            "ms_per_request": ms_graph, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "section 8f rank 4: eval image_encoder -> image_norm -> image_classifier, uint8 "
                                   "3x250x2500 requests", "batch": B, "fused_epilogue": serve.FUSED_EPILOGUE,
                       "launch": "cuda_graph"},
            "eager": {"images_per_s": B / (ms_eager * 1e-3), "ms_per_request": ms_eager, "launches_per_request": launches},
            "with_gradcam": {"images_per_s": B / (ms_cam * 1e-3), "ms_per_request": ms_cam},
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_request": ms_e2e,
                    "h2d_bytes_per_step": host.numel(), "d2h_bytes_per_step": B * 2 * 4},
            "conv_tflops_equiv": round(B * fwd_gflop / ms_graph, 1),
            "cpu_baseline": {"value": n / cpu_s, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{n} images through oracle.model.image_endpoint (fp32, with Grad-CAM)"}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
