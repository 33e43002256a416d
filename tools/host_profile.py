"""Where does the HOST time of one training step go?  (development helper, needs a B200)

    python tools/host_profile.py [--batch 8] [--steps 6]

Runs the bench.py training step at a small batch (so the GPU is never the bottleneck) under cProfile and
prints the functions with the largest own time, plus the wall-clock enqueue time per step."""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--top", type=int, default=45)
    args = ap.parse_args()
    import torch

    import bench
    import ecgmm
    from ecgmm import lib
    from ecgmm import nn as enn
    from ecgmm import optim as eoptim

    lib.require_device()
    dev = torch.device("cuda", 0)

    class Cfg:
        num_classes = 2
        device = dev

    torch.manual_seed(42)
    model = ecgmm.ECGMultimodalModel(Cfg)
    model.train()
    crit = enn.CrossEntropyLoss()
    opt = eoptim.Adam(model.parameters(), lr=1e-4)
    batch = [t.to(dev) for t in bench.synth_batch(args.batch, 42)]

    def step():
        image, ecg, clin, labels = batch
        opt.zero_grad()
        out = model(image, ecg, clin)
        loss = crit(out[3], labels) + 0.1 * out[4]
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0) / args.steps:.2f} ms/step, drained after {1e3 * (t2 - t1):.2f} ms more "
          f"(batch {args.batch})")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.steps):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(args.top)


if __name__ == "__main__":
    main()
