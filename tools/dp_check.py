"""Data-parallel parity check, run under torchrun with N >= 2 GPUs (tests/test_dp_gpu.py drives it):
every rank trains one step on its shard through ecgmm.parallel.DataParallel; the all-reduced
gradient must equal the mean of the per-shard gradients computed WITHOUT communication, weights
must stay identical across ranks after the Adam step, BatchNorm statistics must stay per-rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ecgmm  # noqa: E402
from ecgmm import nn as enn  # noqa: E402
from ecgmm import optim as eoptim  # noqa: E402
from ecgmm.parallel import DataParallel, shard_batch  # noqa: E402
from golden_util import make_inputs, set_dropout  # noqa: E402


def grads_of(model, net, batch, crit):
    model.zero_grad(set_to_none=True)
    out = net(*batch[:3])
    (crit(out[3], batch[3]) + 0.1 * out[4]).backward()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    class Cfg:
        num_classes = 2
        device = dev

    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
    model = ecgmm.ECGMultimodalModel(Cfg)
    set_dropout(model, 0.0)
    model.train()
    dp = DataParallel(model)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    crit = enn.CrossEntropyLoss()
    B = 4 * world
    full = [t.to(dev) for t in make_inputs(5, B, 64, 160, 600)]
    mine = shard_batch(full, rank, world)

    # reference: every shard locally, no communication (restoring BN buffers between runs)
    ref_sum = None
    for r in range(world):
        model.load_state_dict(sd0)
        with dp.no_sync():
            g = grads_of(model, model, shard_batch(full, r, world), crit)
        ref_sum = g if ref_sum is None else {k: ref_sum[k] + g[k] for k in g}
    ref = {k: v / world for k, v in ref_sum.items()}
    model.load_state_dict(sd0)

    got = grads_of(model, dp, mine, crit)
    torch.cuda.synchronize()
    worst = 0.0
    for k, r in ref.items():
        e = float((got[k] - r).norm() / (r.norm() + 1e-20))
        # conv weight gradients are accumulated with fp32 atomics: order-dependent rounding only
        worst = max(worst, e)
    ok = worst < 2e-3 and dp.buckets_last_step >= 7
    eoptim.Adam(model.parameters(), lr=1e-3).step()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_weights = bool(torch.equal(lo, hi))
    # BatchNorm statistics stay per-rank: the running means of different shards must differ
    rm = model.image_encoder.bn1.running_mean.detach().clone()
    rm_lo, rm_hi = rm.clone(), rm.clone()
    dist.all_reduce(rm_lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(rm_hi, op=dist.ReduceOp.MAX)
    ok = ok and not bool(torch.equal(rm_lo, rm_hi))
    flag = torch.tensor([int(ok and same_weights)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DP_CHECK world={world} worst_rel={worst:.3e} buckets={dp.buckets_last_step} "
              f"bytes={dp.bytes_last_step} same_weights={same_weights} result={'OK' if int(flag) else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
