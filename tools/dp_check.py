"""Data-parallel parity check, run under torchrun (tests/test_dp_gpu.py drives it): every rank trains one step on its
shard through ecgmm.parallel.DataParallel.  Checked:
  * the all-reduced gradient equals the mean of the per-shard gradients computed WITHOUT communication (the kernels are
    deterministic, so this holds to fp32 summation order: 1e-5);
  * it is the mean over shards of the ORACLE's per-shard gradients (each rank runs the fp32 CPU oracle and its
    bf16-emulated variant on its own shard; the means are formed with all_reduce; tolerance as in tests/parity_util.py);
  * weights stay identical across ranks after the Adam step; BatchNorm statistics stay per-rank.
N >= 2 GPUs: NCCL, one rank per GPU.  DP_CHECK_ONE_GPU=1: the ranks share GPU 0 and communicate through gloo (the host
logic and the arena aliasing are the same; this is what a 1-GPU test box can run)."""
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
    one_gpu = os.environ.get("DP_CHECK_ONE_GPU", "0") == "1"
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
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
    scale = max(float(r.norm()) for r in ref.values())
    for k, r in ref.items():
        if float(r.norm()) < 1e-6 * scale:
            continue  # Conv1d biases: true gradient 0
        worst = max(worst, float((got[k] - r).norm() / r.norm()))
    ok = worst < 1e-5 and dp.buckets_last_step >= 7

    # ---- the same averaged gradient against the oracle: mean over shards of the fp32 CPU gradients
    from oracle import model as om
    from parity_util import GRAD_FLOOR, GRAD_MEDIAN_X, GRAD_NOISE_X, GRAD_REL, bf16_emulated_oracle

    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    ora = om.ECGMultimodalModel()
    ora.load_state_dict({k: v.cpu() for k, v in sd0.items()})
    set_dropout(ora, 0.0)
    ora.train()
    cpu_shard = [t.cpu() for t in mine]
    emul = bf16_emulated_oracle(ora, *cpu_shard)
    o_out = ora(*cpu_shard[:3])
    om.fusion_loss(o_out, cpu_shard[3]).backward()
    names = [k for k, p in ora.named_parameters() if p.grad is not None]
    o_flat = torch.cat([dict(ora.named_parameters())[k].grad.flatten() for k in names]).to(dev)
    e_flat = torch.cat([emul[k].flatten() for k in names]).to(dev)
    for t in (o_flat, e_flat):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    sizes = [dict(ora.named_parameters())[k].numel() for k in names]
    o_mean = dict(zip(names, o_flat.split(sizes)))
    e_mean = dict(zip(names, e_flat.split(sizes)))
    oscale = max(float(v.norm()) for v in o_mean.values())
    live = [k for k in names if float(o_mean[k].norm()) >= GRAD_FLOOR * oscale]
    e_rel = {k: float((e_mean[k] - o_mean[k]).norm() / o_mean[k].norm()) for k in live}
    floor = GRAD_MEDIAN_X * sorted(e_rel.values())[len(e_rel) // 2]
    worst_vs_oracle, bad = 0.0, []
    for k in live:
        rel = float((got[k].flatten() - o_mean[k]).norm() / o_mean[k].norm())
        allowed = max(GRAD_REL, GRAD_NOISE_X * e_rel[k], floor)
        worst_vs_oracle = max(worst_vs_oracle, rel / allowed)
        if rel > allowed:
            bad.append((k, rel, allowed))
    ok = ok and not bad
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
        print(f"DP_CHECK world={world} backend={dist.get_backend()} worst_rel_vs_local_mean={worst:.3e} "
              f"worst_vs_oracle_allowed={worst_vs_oracle:.3f} oracle_failures={bad[:3]} buckets={dp.buckets_last_step} "
              f"bytes={dp.bytes_last_step} same_weights={same_weights} result={'OK' if int(flag) else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
