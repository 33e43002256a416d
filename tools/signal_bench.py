"""configs[1] on one B200: signal_model.py's 12-lead ResNet1D-SE (12 x 5000 @ 500 Hz, batch 256) -- FocalLoss + Adam
training step, plus the batched GPU preprocessing (baseline removal + Butterworth filtfilt + z-score) that the
reference runs per sample on the host (signal_model.py:203-224).  Development / profiles helper.

    python tools/signal_bench.py [--batch 256] [--length 5000] [--steps 10] [--cpu-steps 2]      (= bench.py --config signal)

One JSON line: train samples/s (device-resident inputs, CUDA events; the step replayed as one CUDA graph, the eager
kernel-by-kernel loop next to it), every kernel class of one step against BOTH rooflines (the 1-D convolutions sit far
below the ridge: their HBM figure is the one that matters), preprocessing signals/s and
its share, and the fp32 oracle (CPU) on a bounded sample for both."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--length", type=int, default=5000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args(argv)
    import numpy as np
    import torch

    import ecgmm
    from ecgmm import lib, ops, preprocess
    from ecgmm.graph import GraphedTrainStep
    from ecgmm import nn as enn
    from ecgmm import optim as eoptim
    from oracle import model as om
    from oracle import preprocess as op

    lib.require_device()
    dev = torch.device("cuda", 0)
    B, L = args.batch, args.length
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, 12, L, generator=g)
    y = torch.randint(0, 2, (B,), generator=g)
    xd, yd = x.to(dev), y.to(dev)
    torch.manual_seed(42)
    net = ecgmm.ResNet1D_SE(12, 2).to(dev).train()
    crit = enn.FocalLoss()
    opt = eoptim.Adam(net.parameters(), lr=1e-3)

    def step(inp):
        opt.zero_grad()
        loss = crit(net(inp), yd)
        loss.backward()
        opt.step()
        return loss

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    pre_ms = timed(lambda: preprocess.preprocess_signal(xd, zscore=True), args.steps)
    xp = preprocess.preprocess_signal(xd, zscore=True)
    eager_ms = timed(lambda: step(xp), args.steps)
    # per-kernel-class device times of one eager step (CUDA events on the launching stream)
    ops.PROFILE = []
    n0 = lib.launch_count()
    step(xp)
    torch.cuda.synchronize()
    launches = lib.launch_count() - n0
    prof, ops.PROFILE = ops.PROFILE, None
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks = json.load(open(pp))
    classes = {}
    for kind, work, a, b, nbytes in prof:
        c = classes.setdefault(kind.split("/")[0], {"launches": 0, "ms": 0.0, "work": 0.0, "bytes": 0.0})
        c["launches"] += 1
        c["ms"] += a.elapsed_time(b)
        c["work"] += work
        c["bytes"] += nbytes if nbytes is not None else work  # bandwidth-bound kinds carry bytes as their work
    kernels = {}
    for kind, c in classes.items():
        conv = kind.startswith("conv") or kind.startswith("signal_stem")
        gbs = c["bytes"] / (c["ms"] * 1e-3) / 1e9
        kernels[kind] = {"launches": c["launches"], "ms_per_step": round(c["ms"], 4), "bound": "hbm",
                         "hbm_GBs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4)}
        if conv:
            tf = c["work"] / (c["ms"] * 1e-3) / 1e12
            kernels[kind].update(tflops=round(tf, 1), tensor_frac=round(tf / peaks["bf16_tflops_sustained"], 4))
    timed_ms = sum(c["ms"] for c in classes.values())
    # the step as one CUDA graph (ecgmm.graph.GraphedTrainStep): ~300 launches of a few microseconds each are
    # host-bound when issued one by one
    gstep = GraphedTrainStep(net, crit, opt, [xp, yd], restore=False)
    train_ms = timed(lambda: gstep(xp, yd), args.steps)

    def both():
        return gstep(preprocess.preprocess_signal(xd, zscore=True), yd)

    both_ms = timed(both, args.steps)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = om.ResNet1D_SE(12, 2).train()
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    rcrit = om.FocalLoss()
    nb = min(B, 64)

    def cpu_step():
        ropt.zero_grad()
        loss = rcrit(ref(x[:nb]), y[:nb])
        loss.backward()
        ropt.step()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(args.cpu_steps):
        cpu_step()
    cpu_train = nb * args.cpu_steps / (time.perf_counter() - t0)
    ns = 24
    xn = x.reshape(-1, L)[:ns].numpy().astype(np.float64)
    t0 = time.perf_counter()
    op.preprocess_batch(xn, zscore=True)
    cpu_pre = ns / (time.perf_counter() - t0)
    line = {
        "metric": "train samples/sec, signal_model.py ResNet1D-SE 12x%d, batch %d, 1 B200" % (L, B),
        "value": B / (train_ms * 1e-3), "unit": "samples/s", "ms_per_step": train_ms, "n_gpus": 1, "dtype": "bf16",
        "config": {"workload": "configs[1]: 12-lead 1D-CNN, FocalLoss + Adam", "batch": B, "leads": 12, "length": L,
                   "launch": "cuda_graph"},
        "eager": {"value": B / (eager_ms * 1e-3), "unit": "samples/s", "ms_per_step": eager_ms,
                  "launches_per_step": launches},
        "kernels": kernels, "kernels_ms_sum": round(timed_ms, 4),
        "with_preprocessing": {"value": B / (both_ms * 1e-3), "unit": "samples/s", "ms_per_step": both_ms},
        "preprocessing": {"signals_per_s": B * 12 / (pre_ms * 1e-3), "ms_per_batch": pre_ms,
                          "what": "baseline removal + butter(5) filtfilt + z-score, float64, one launch per batch",
                          "hbm_GBs": (B * 12 * L * 8) / (pre_ms * 1e-3) / 1e9},
        "cpu_baseline": {"train_samples_per_s": cpu_train, "preprocess_signals_per_s": cpu_pre, "cores": cores,
                         "kind": "port", "sample": f"{args.cpu_steps} train steps of batch {nb} through oracle.model.ResNet1D_SE; "
                                                   f"{ns} signals through oracle.preprocess (numpy/scipy, 1 thread)"},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
