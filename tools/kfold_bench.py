"""configs[4]: train_kfold.py-style nested cross-validation (train_kfold.py:135-178: 5 outer folds, 3 inner folds each =
15 independent training jobs + the outer test of the best inner model) with the jobs sharded across the GPUs of one
box, on a synthetic tri-modal data set.

    python tools/kfold_bench.py [--patients 10000] [--folds 5] [--epochs 1] [--batch 64] [--height 224 --width 224]
    (= bench.py --config kfold)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/kfold_bench.py

Training jobs are independent: job j = (outer fold, inner fold) runs on rank j mod world
(ecgmm.parallel.folds_for_rank), no collective on the data path; every job also scores its model on its outer fold's
test split, and rank 0 gathers times and accuracies at the end and reports, per outer fold, the test accuracy of the
inner model with the best validation accuracy (train_kfold.py:170-175 loads "best_inner.pth" for that).
StratifiedKFold(k, shuffle, seed 42) for both levels as in train_kfold.py:137,151.  The data set is synthetic and SEPARABLE (the label shifts the clinical features and the signal
amplitude), so the held-out accuracy shows that the folds really train.  Image size: 224 x 224 (SURVEY.md section 8d
allows "the same per-sample shapes as cfg3 or 224^2 for tractability -- state which": 10 000 patients are 1.5 GB of
uint8 pixels on the host at 224^2 and 18.75 GB at 250 x 2500; pass --height 250 --width 2500 for the native size).
Each fold uses ecgmm.graph.GraphedTrainStep (one launch per step) with fusion_only=True (the single-tensor API of
train_kfold.py:59-64)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--patients", type=int, default=10000)
    ap.add_argument("--folds", type=int, default=5, help="outer folds (config.py k_outer)")
    ap.add_argument("--inner", type=int, default=3, help="inner folds per outer fold (config.py k_inner); 0: plain k-fold")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--height", type=int, default=224)
    ap.add_argument("--width", type=int, default=224)
    ap.add_argument("--length", type=int, default=2476)
    args = ap.parse_args(argv)
    import numpy as np
    import torch
    import torch.distributed as dist
    from sklearn.model_selection import StratifiedKFold

    import ecgmm
    from ecgmm import lib
    from ecgmm import nn as enn
    from ecgmm import optim as eoptim
    from ecgmm.graph import GraphedTrainStep
    from ecgmm.parallel import folds_for_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib.require_device()
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(42)
    P = args.patients
    labels = (torch.rand(P, generator=g) < 0.4).long()
    images = torch.randint(0, 256, (P, 3, args.height, args.width), generator=g, dtype=torch.uint8)  # raw pixels
    sig = torch.randn(P, args.length, generator=g) * (1.0 + 0.5 * labels.float()).unsqueeze(1)
    clin = torch.randn(P, 24, generator=g) + 0.8 * labels.float().unsqueeze(1)
    skf = StratifiedKFold(n_splits=args.folds, shuffle=True, random_state=42)
    lab_np = labels.numpy()
    jobs = []  # (outer fold, inner fold or -1, train indices, validation indices or None, outer test indices)
    for o, (train_val, test) in enumerate(skf.split(np.arange(P), lab_np)):
        if args.inner > 1:
            inner = StratifiedKFold(n_splits=args.inner, shuffle=True, random_state=42)
            for i, (itr, iva) in enumerate(inner.split(train_val, lab_np[train_val])):
                jobs.append((o, i, train_val[itr], train_val[iva], test))
        else:
            jobs.append((o, -1, train_val, None, test))
    results = []
    for k in folds_for_rank(len(jobs), rank, world):
        o_fold, i_fold, tr, va, te = jobs[k]
        tr = torch.from_numpy(tr)
        te = torch.from_numpy(te)
        va = torch.from_numpy(va) if va is not None else None

        class Cfg:
            num_classes = 2
            device = dev

        torch.manual_seed(42 + k)
        model = ecgmm.ECGMultimodalModel(Cfg, fusion_only=True).train()
        crit = enn.CrossEntropyLoss()
        opt = eoptim.Adam(model.parameters(), lr=1e-3)
        B = args.batch

        def batch_of(idx):
            return [images[idx].to(dev, non_blocking=True), sig[idx].to(dev), clin[idx].to(dev), labels[idx].to(dev)]

        step = GraphedTrainStep(model, crit, opt, batch_of(tr[:B]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seen = 0
        for ep in range(args.epochs):
            perm = tr[torch.randperm(len(tr), generator=g)]
            for i in range(0, len(perm) - B + 1, B):
                step(*batch_of(perm[i:i + B]))
                seen += B
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        model.eval()

        def accuracy(split):
            correct = 0
            with torch.no_grad():
                for i in range(0, len(split), 256):
                    idx = split[i:i + 256]
                    out = model(images[idx].to(dev), sig[idx].to(dev), clin[idx].to(dev))
                    correct += int((out.argmax(1).cpu() == labels[idx]).sum())
            return round(correct / max(1, len(split)), 4)

        results.append({"job": k, "fold": o_fold, "inner": i_fold, "rank": rank, "train_samples": seen,
                        "seconds": round(dt, 3), "samples_per_s": round(seen / dt, 1),
                        "val_acc": accuracy(va) if va is not None else None, "heldout_acc": accuracy(te)})
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
        dist.destroy_process_group()
    if rank == 0:
        results.sort(key=lambda r: r["job"])
        outer = []
        for o in range(args.folds):
            mine = [r for r in results if r["fold"] == o]
            best = max(mine, key=lambda r: (r["val_acc"] if r["val_acc"] is not None else r["heldout_acc"]))
            outer.append({"fold": o, "best_inner": best["inner"], "test_acc": best["heldout_acc"]})
        per_rank = {}
        for r in results:
            per_rank[r["rank"]] = per_rank.get(r["rank"], 0.0) + r["seconds"]
        print(json.dumps({"metric": "k-fold training, folds sharded over GPUs", "n_gpus": world, "unit": "samples/s",
                          "value": round(sum(r["train_samples"] for r in results) / max(per_rank.values()), 1),
                          "higher_is_better": True,
                          "config": {"workload": "configs[4]: nested CV, training jobs sharded over GPUs", "patients": P,
                                     "folds": args.folds, "inner_folds": args.inner, "jobs": len(jobs), "epochs": args.epochs,
                                     "image": [3, args.height, args.width], "batch": args.batch, "input": "uint8 pixels"},
                          "wall_s_training_max_over_ranks": round(max(per_rank.values()), 3),
                          "samples_per_s_sum": round(sum(r["samples_per_s"] for r in results), 1),
                          "outer_folds": outer, "jobs": results}))


if __name__ == "__main__":
    main()
